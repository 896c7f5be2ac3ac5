"""miniweatherml_b200: B200-native (sm_100a) implementation of miniWeatherML's time-stepping hot path.

The product is libmwb200.so (hand-written CUDA behind the C ABI of include/mw_b200.h) plus the C++ host
headers under miniweatherml_b200/host/ that mirror the reference's module interface.  This Python package
only binds the C ABI with ctypes for the tests and bench.py; PyTorch is used for device memory and
torch.distributed plumbing, nothing else.  There is no CPU fallback: without the built extension and a
B200 every compute call raises.
"""
from .capi import (MwError, Config, Dycore, lib, lib_path, make_config, weno5_edges, kessler_step,  # noqa: F401
                   mlp_forward, surrogate_forward, sponge_layer, column_average, nudge_to_column,
                   perturb_temperature, city_layout, extract_column, horizontal_sponge_apply,
                   time_average_accumulate, mlp_dense2_forward, mean_difference, probe_fp64_rate)
