"""Reference flashinfer/tllm_utils.py: ``delay_kernel`` (spin on the GPU for a number of microseconds; used by tests that need
work in flight on a stream)."""
import torch


def delay_kernel(stream_delay_micro_secs: int) -> None:
    if torch.cuda.is_available():
        torch.cuda._sleep(int(stream_delay_micro_secs * 1.9e3))  # cycles at ~1.9 GHz


def get_trtllm_utils_module():
    import sys

    return sys.modules[__name__]
