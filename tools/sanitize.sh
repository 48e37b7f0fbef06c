#!/bin/bash
# compute-sanitizer passes over a small, fast subset of the GPU tests (run on a B200 box, e.g. through gpurun):
#   tools/sanitize.sh memcheck    # out-of-bounds / misaligned accesses, leaked allocations
#   tools/sanitize.sh racecheck   # shared-memory data races (mbarrier / named-barrier protocols of the tcgen05 kernels)
#   tools/sanitize.sh synccheck   # illegal barrier usage (divergent bar.sync / cluster barriers)
#   tools/sanitize.sh initcheck   # reads of uninitialised global memory (MoE padding rows are expected to show up here)
# The reference runs no sanitizer in CI; these are the manual gates used while developing the kernels.
set -u
tool=${1:-memcheck}
shift || true
tests=${*:-"tests/test_gpu_elementwise.py tests/test_gpu_quant_topk.py tests/test_gpu_mla.py::test_mla_decode_basic"}
cd "$(dirname "$0")/.."
exec compute-sanitizer --tool "$tool" --error-exitcode 1 --kernel-name-exclude regex:at:: \
  python -m pytest $tests -x -q -m gpu -p no:cacheprovider
