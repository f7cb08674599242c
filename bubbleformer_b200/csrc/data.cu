// Input pipeline on the device: the trajectories live in HBM as one (frames, C, H*W) fp32
// tensor, and a batch of forecasting windows is cut, normalised and laid out (B, T, C', H, W) by one kernel --
// upstream data/dataset.py:120-186 (BubbleForecast.__getitem__: per-field slicing, (x - diff) / div, stack, permute)
// run per sample on the host through h5py.
//   out[b, t, j, :] = (frames[first[b] + t_off + t, ch[j], :] - diff[ch[j]]) * inv_div[ch[j]]
#include "common.cuh"

namespace bf {

struct WindowArgs {
  const float* frames; const long* first; const int* ch; const float* diff; const float* inv_div;
  float* out;
  int T, C_src, C_out, t_off;
  long HW4;          // float4 per field
};

__global__ void __launch_bounds__(256) window_gather_kernel(WindowArgs a) {
  pdl_prologue_done();
  const int b = blockIdx.y;
  const long per_sample = (long)a.T * a.C_out * a.HW4;
  const float4* src = reinterpret_cast<const float4*>(a.frames) + (a.first[b] + a.t_off) * a.C_src * a.HW4;
  float4* dst = reinterpret_cast<float4*>(a.out) + (long)b * per_sample;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += (long)gridDim.x * blockDim.x) {
    const long plane = i / a.HW4, px = i - plane * a.HW4;
    const int t = (int)(plane / a.C_out), j = (int)(plane - (long)t * a.C_out);
    const int c = a.ch[j];
    const float d = a.diff[c], s = a.inv_div[c];
    const float4 v = __ldg(src + ((long)t * a.C_src + c) * a.HW4 + px);
    dst[i] = make_float4((v.x - d) * s, (v.y - d) * s, (v.z - d) * s, (v.w - d) * s);
  }
}

}  // namespace bf

using namespace bf;

extern "C" int bf_window_gather(const float* frames, const int64_t* first_frame, const int32_t* channels, const float* diff,
                                const float* inv_div, float* out, int B, int T, int C_src, int C_out, int64_t HW, int t_off,
                                void* stream) {
  BF_REQUIRE(frames && first_frame && channels && diff && inv_div && out, "bf_window_gather: null pointer");
  BF_REQUIRE(B > 0 && B <= 65535 && T > 0 && C_src > 0 && C_out > 0 && HW > 0 && HW % 4 == 0 && t_off >= 0,
             "bf_window_gather: bad geometry (H*W must be a multiple of 4)");
  BF_REQUIRE(((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "bf_window_gather: alignment");
  WindowArgs a{frames, reinterpret_cast<const long*>(first_frame), channels, diff, inv_div, out, T, C_src, C_out, t_off, HW / 4};
  const long per_sample = (long)T * C_out * (HW / 4);
  long bx = (per_sample + 255) / 256;
  const long cap = (8L * num_sms() + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  launch_k(window_gather_kernel, dim3((unsigned)bx, (unsigned)B), dim3(256), (size_t)0, static_cast<cudaStream_t>(stream), a);
  count_launch();
  BF_LAUNCH_CHECK("window_gather_kernel");
  return BF_OK;
}
