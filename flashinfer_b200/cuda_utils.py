"""Reference flashinfer/cuda_utils.py: unwrap cuda-python ``(err, *values)`` results."""


def checkCudaErrors(result):
    err = result[0]
    if int(err) != 0:
        name = getattr(err, "name", str(err))
        raise RuntimeError(f"CUDA error code={int(err)} ({name})")
    if len(result) == 1:
        return None
    return result[1] if len(result) == 2 else result[1:]
