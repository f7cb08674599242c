"""`bubbleformer.data.BubbleForecast` (upstream data/dataset.py:16-186) backed by HBM-resident trajectories.

Same constructor arguments and the same `__len__`, `normalize`, `__getitem__` contract;
`__getitem__` returns CUDA tensors, and `batch(indices)` builds a whole batch with one kernel per window."""
from bubbleformer_b200.data import DeviceForecastWindows


class BubbleForecast(DeviceForecastWindows):
    def __init__(self, filenames, input_fields=None, output_fields=None, norm="none", downsample_factor=1, time_window=16,
                 start_time=50, return_fluid_params=False):
        super().__init__(filenames, input_fields=input_fields, output_fields=output_fields, norm=norm,
                         time_window=time_window, start_time=start_time, return_fluid_params=return_fluid_params,
                         downsample_factor=downsample_factor)
