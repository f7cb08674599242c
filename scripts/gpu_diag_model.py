"""Whole-model parity diagnostics on the B200 box against the golden fixtures (reference outputs, fp64).

    python scripts/gpu_diag_model.py [case ...]      # default: all cases -> gpurun_out/diag_model.log
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["film_eval_e128", "film_train_masks_e128", "avit_generic_e96", "rollout", "shapes", "big512",
         "oracle_cfg2", "oracle_cfg5_512", "oracle_cfg5_1024", "oracle_cfg4_strip"]

SMALL = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=384, num_heads=6, processor_blocks=12,
             drop_path=0.2, attn_scale=True, feat_scale=True, num_fluid_params=9)      # film_avit_small.yaml
BIG = dict(SMALL, embed_dim=768, num_heads=12)                                          # film_avit_big.yaml


def _host_mem_gb():
    """Memory this process may still use on the host (cgroup limit when there is one)."""
    import psutil
    avail = psutil.virtual_memory().available
    for f in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(f).read().strip()
            if v.isdigit():
                used = 0
                for u in ("/sys/fs/cgroup/memory.current", "/sys/fs/cgroup/memory/memory.usage_in_bytes"):
                    try:
                        used = int(open(u).read().strip())
                        break
                    except OSError:
                        pass
                avail = min(avail, int(v) - used)
        except OSError:
            pass
    return avail / 2**30


def oracle_case(name, cfg, B, T, H, W, seed, train=True):
    """Whole-model fwd + loss + every gradient on the CUDA path vs the CPU oracle on the host, same weights / inputs /
    stochastic-depth masks, at a BASELINE.json shape (the code path bench.py times)."""
    import torch
    from bubbleformer_b200 import get_model
    from bubbleformer_b200.losses import rel_l2_loss
    from oracle import parity
    case = parity.make_case(cfg, B, T, H, W, seed, train=train)
    ref = parity.oracle_run(case["sd"], case["x"], case["tgt"], case["cond"], cfg, case["masks"])
    print(f"[{name}] oracle on {torch.get_num_threads()} host threads: {ref['seconds']:.1f} s "
          f"(E={cfg['embed_dim']} blocks={cfg['processor_blocks']} B={B} T={T} {H}x{W})", flush=True)
    model = get_model("filmavit", time_window=T, **cfg).to("cuda")
    model.load_state_dict(case["sd"], strict=True)
    got = parity.candidate_run(model, case["x"], case["tgt"], case["cond"], case["masks"], loss_fn=rel_l2_loss)
    res = parity.compare(ref, got, verbose_prefix=name)
    return parity.passes(res)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def run_case(name, exact=False):
    """exact=True: the fp32 validation configuration (bubbleformer_b200.set_exact_mode) at the north star's 1e-4."""
    if exact:
        import bubbleformer_b200
        bubbleformer_b200.set_exact_mode(True)
        try:
            return _run_case(name, fwd_tol=1e-4, grad_tol=1e-4)
        finally:
            bubbleformer_b200.set_exact_mode(False)
    return _run_case(name)


def _run_case(name, fwd_tol=1e-2, grad_tol=2e-2):
    import json
    import numpy as np
    import torch
    from bubbleformer_b200 import get_model
    from oracle import filmavit_oracle as O
    from tests.helpers import load_case, GOLD
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    ok = True
    if name in ("film_eval_e128", "film_train_masks_e128", "avit_generic_e96"):
        case = load_case(name, dtype=torch.float32)
        cfg = dict(case["cfg"])
        model = get_model(case["model"], time_window=case["T"], **cfg).to(dev)
        model.load_state_dict(case["sd"], strict=True)
        if case["masks"] is not None:
            model.train()
            model.drop_masks_override = [tuple(None if m is None else m.to(dev) for m in trip) for trip in case["masks"]]
        else:
            model.eval()
        x = case["x"].to(dev).requires_grad_(True)
        tgt = case["tgt"].to(dev)
        args = (x, case["cond"].to(dev)) if case["cond"] is not None else (x,)
        y = model(*args)
        gold = case["gold"]
        yref = torch.from_numpy(gold["y"]).to(dev)
        for c in range(y.shape[2]):
            e = rel(y[:, :, c], yref[:, :, c])
            print(f"[{name}] fwd channel {c}: rel-L2 {e:.3e}")
            ok &= e < fwd_tol
        loss = O.rel_l2_loss(y, tgt)
        print(f"[{name}] loss {float(loss):.6f} ref {float(gold['loss']):.6f}")
        loss.backward()
        e = rel(x.grad, torch.from_numpy(gold["dx"]).to(dev))
        print(f"[{name}] dx rel-L2 {e:.3e}")
        ok &= e < grad_tol
        gn = np.sqrt(sum(float((gold["grad/" + k].astype(np.float64) ** 2).sum()) for k, _ in model.named_parameters()))
        tot = 0.0
        worst = []
        for k, p in model.named_parameters():
            ref = torch.from_numpy(gold["grad/" + k]).to(dev)
            if p.grad is None:
                print(f"   {k}: NO GRAD")
                ok = False
                continue
            d = float((p.grad.double() - ref.double()).norm())
            tot += d * d
            r = d / max(float(ref.double().norm()), 1e-30)
            worst.append((r, d / gn, k, float(ref.norm())))
        worst.sort(reverse=True)
        for r, gr, k, rn in worst[:25]:
            print(f"   {k:60s} rel {r:.3e}  global-rel {gr:.3e}  |ref| {rn:.3e}")
        ge = np.sqrt(tot) / gn
        print(f"[{name}] all parameter gradients: global-norm-relative error {ge:.3e}")
        ok &= ge < grad_tol
    elif name == "rollout":
        z = np.load(os.path.join(GOLD, "rollout_sample1_small.npz"))
        meta = json.loads(str(z["meta"]))
        from oracle.param_init import fluid_params, param_shapes, random_state_dict
        cfg = meta["cfg"]
        shapes = param_shapes(**{k: v for k, v in cfg.items() if k != "drop_path"})
        sd = random_state_dict(shapes, seed=meta["seed"], dtype=torch.float32)
        model = get_model("filmavit", time_window=5, **cfg).to(dev)
        model.load_state_dict(sd, strict=True)
        model.eval()
        cond = fluid_params(1, torch.float32).to(dev)
        preds = torch.from_numpy(z["preds"]).to(dev)
        inp0 = torch.from_numpy(z["inp0"]).to(dev).unsqueeze(0)
        with torch.no_grad():
            free = inp0
            for s in range(preds.shape[0]):
                tf_in = inp0 if s == 0 else preds[s - 1].unsqueeze(0)
                y = model(tf_in, cond)
                errs = [rel(y[0, :, c], preds[s][:, c]) for c in range(4)]
                free = model(free, cond)
                ferr = [rel(free[0, :, c], preds[s][:, c]) for c in range(4)]
                env = z["envelope"][s]
                print(f"[rollout] step {s+1}: teacher-forced " + " ".join(f"{e:.2e}" for e in errs)
                      + " | free-running " + " ".join(f"{e:.2e}" for e in ferr)
                      + " | ref chaos envelope " + " ".join(f"{e:.2e}" for e in env))
                ok &= max(errs) < fwd_tol
    elif name == "shapes":
        # the upstream shape tests (models/tests/test_get_model.py, layers/tests/test_patching.py), a subset
        from bubbleformer_b200.layers import HMLPDebed, HMLPEmbed
        for patch in (4, 8, 16, 32):
            for E in (192, 384, 768, 1024):
                emb = HMLPEmbed(patch_size=patch, in_channels=4, embed_dim=E).to(dev)
                deb = HMLPDebed(patch_size=patch, out_channels=4, embed_dim=E).to(dev)
                x = torch.randn(1, 4, 64, 64, device=dev)
                with torch.no_grad():
                    yy = emb(x)
                    zz = deb(yy)
                good = yy.shape == (1, E, 64 // patch, 64 // patch) and zz.shape == x.shape and bool(torch.isfinite(zz).all())
                print(f"[shapes] patch {patch} E {E}: {tuple(yy.shape)} {tuple(zz.shape)} {'OK' if good else 'BAD'}")
                ok &= good
        for (fi, fo, patch, E, a_s, f_s) in [(1, 2, 8, 192, True, False), (2, 1, 16, 384, False, True), (2, 2, 8, 384, True, True)]:
            m = get_model("avit", input_fields=fi, output_fields=fo, time_window=3, patch_size=patch, embed_dim=E,
                          num_heads=4, processor_blocks=4, drop_path=0.1, attn_scale=a_s, feat_scale=f_s).to(dev)
            x = torch.randn(2, 3, fi, 64, 64, device=dev)
            y = m(x)
            y.square().mean().backward()
            good = y.shape == (2, 3, fo, 64, 64) and bool(torch.isfinite(y).all()) and all(
                p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
            print(f"[shapes] avit fi={fi} fo={fo} patch={patch} E={E} attn_scale={a_s} feat_scale={f_s}: {'OK' if good else 'BAD'}")
            ok &= good
    elif name == "oracle_cfg2":
        # BASELINE configs[1]: film_avit_small, T=5, 4 fields, 512x512 (h=w=32, P=1024), train mode with masks, B=2
        ok &= oracle_case(name, SMALL, 2, 5, 512, 512, seed=21)
    elif name == "oracle_cfg5_512":
        ok &= oracle_case(name, BIG, 1, 5, 512, 512, seed=22)
    elif name == "oracle_cfg5_1024":
        # BASELINE configs[4]: film_avit_big at 1024x1024 (h=w=64, L=64).  The fp32 autograd graph of the host oracle needs
        # ~6.5 GB per block at this size; the depth is reduced when the host cannot hold all 12 (every block runs the same
        # kernels, the full depth is checked at 512x512 above).
        mem = _host_mem_gb()
        blocks = 12 if mem > 110 else max(2, min(12, int((mem - 12) / 7)))
        print(f"[{name}] host memory available {mem:.0f} GiB -> {blocks} blocks", flush=True)
        ok &= oracle_case(name, dict(BIG, processor_blocks=blocks), 1, 5, 1024, 1024, seed=23)
    elif name == "oracle_cfg4_strip":
        # BASELINE configs[3]: non-square flow-boiling strip 128x1024 (h=8, w=64): 3 teacher-forced rollout steps
        # (eval, no_grad; the input of step k is the ORACLE's output of step k-1) + one fwd/bwd parity at the same shape
        from bubbleformer_b200 import get_model
        from oracle import parity
        cfg = dict(SMALL, drop_path=0.0)
        case = parity.make_case(cfg, 1, 5, 128, 1024, seed=24, train=False)
        model = get_model("filmavit", time_window=5, **cfg).to(dev).eval()
        model.load_state_dict(case["sd"], strict=True)
        inp = case["x"]
        for s in range(3):
            ref = parity.oracle_run(case["sd"], inp, None, case["cond"], cfg, None, grads=False)
            got = parity.candidate_run(model, inp, None, case["cond"], None, grads=False)
            res = parity.compare(ref, got, verbose_prefix=f"{name} step {s + 1}")
            ok &= parity.passes(res)
            inp = ref["y"]
        ok &= oracle_case(name + " fwd+bwd", SMALL, 1, 5, 128, 1024, seed=25)
    elif name == "big512":
        # config 2 shape: does it run, how long does it take, how much memory
        cfg = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=384, num_heads=6, processor_blocks=12,
                   drop_path=0.2, attn_scale=True, feat_scale=True, num_fluid_params=9)
        from oracle.param_init import fluid_params
        m = get_model("filmavit", time_window=5, **cfg).to(dev)
        m.train()
        x = torch.randn(8, 5, 4, 512, 512, device=dev)
        tgt = torch.randn(8, 5, 4, 512, 512, device=dev)
        cond = fluid_params(8).to(dev)
        from bubbleformer_b200 import _lib
        for it in range(4):
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            n0 = _lib.launch_count()
            e0.record()
            y = m(x, cond)
            e1.record()
            loss = O.rel_l2_loss(y, tgt)
            loss.backward()
            e2.record()
            torch.cuda.synchronize()
            print(f"[big512] iter {it}: fwd {e0.elapsed_time(e1):.2f} ms  bwd(+loss) {e1.elapsed_time(e2):.2f} ms  "
                  f"launches {_lib.launch_count() - n0}  loss {float(loss):.4f}  "
                  f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
            for p in m.parameters():
                p.grad = None
        ok &= bool(torch.isfinite(y).all())
    else:
        raise SystemExit(f"unknown case {name}")
    torch.cuda.synchronize()
    return ok


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        ok = all([run_case(c) for c in sys.argv[1:]])
        sys.exit(0 if ok else 3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "diag_model.log"), "w")
    summary = []
    for c in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True, timeout=420)
            out, code = r.stdout + r.stderr, r.returncode
        except subprocess.TimeoutExpired as e:
            out, code = (e.stdout or b"").decode() + (e.stderr or b"").decode() + "\nTIMEOUT", -9
        log.write(f"==== {c} (exit {code})\n{out}\n")
        log.flush()
        print(f"==== {c} (exit {code})\n" + "\n".join(out.strip().splitlines()[-60:]), flush=True)
        summary.append((c, code))
    s = "SUMMARY " + " ".join(f"{c}:{k}" for c, k in summary)
    print(s)
    log.write(s + "\n")


if __name__ == "__main__":
    main()
