// POD-Attention as ONE kernel: prefill and decode CTAs in a single grid.
//
// Parity: reference include/flashinfer/attention/pod.cuh + flashinfer/pod.py (PODWithPagedKVCacheWrapper): one launch whose
// CTAs run either the prefill or the decode program so that the compute-bound prefill tiles and the bandwidth-bound decode
// tiles overlap on the same GPU.
//
// B200-first design: both attention programs are already persistent, work-list driven tcgen05 kernels (one CTA per SM,
// each CTA walks its planner-assigned slice), so fusing them needs no SM-id tricks: the host splits the 148 SMs with a
// cost model (prefill FLOPs vs. decode bytes), plans each side for its budget, and this kernel dispatches on the CTA index:
// CTAs [0, n_prefill) run prefill_body (prefill_sm100.cu), the rest run decode_body (decode_sm100.cu).  The two source
// files are compiled into this translation unit inside namespaces; their host launchers are reused verbatim: with the
// stage armed (pod_arm), prefill_run parks its fully built launch (tensor maps + parameter block) instead of launching,
// and the decode launcher that follows picks it up and launches the fused kernel.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>
#include <cuda_fp8.h>
#include <cstring>
#include <type_traits>

#define FIB_POD_TU 1

namespace fib200 {
struct PodStage {
  bool armed = false, have_prefill = false;
  CUtensorMap tm[3];
  alignas(16) unsigned char params[512];
  uint32_t idesc[2];
  int grid = 0, smem = 0, f16 = 0, dqk = 0;
};
inline PodStage& pod_stage() {
  static thread_local PodStage s;
  return s;
}
// defined below, once both programs are visible
template <int NV, typename T>
int pod_launch(const CUtensorMap& dK, const CUtensorMap& dV, const void* dparams, size_t dparams_size, uint32_t di_qk,
               uint32_t di_pv, int dgrid, int dsmem, bool pdl, cudaStream_t stream);
}  // namespace fib200

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace podp {
#include "prefill_sm100.cu"
}
namespace podd {
#include "decode_sm100.cu"
}

namespace {

template <typename T, int NV>
__global__ void __launch_bounds__(384, 1)
pod_kernel(const __grid_constant__ CUtensorMap pQ, const __grid_constant__ CUtensorMap pK, const __grid_constant__ CUtensorMap pV,
           const podp::PrefillParams pp, uint32_t pi_qk, uint32_t pi_pv, const __grid_constant__ CUtensorMap dK,
           const __grid_constant__ CUtensorMap dV, const podd::DecodeParams dp, uint32_t di_qk, uint32_t di_pv, int n_prefill) {
  if (int(blockIdx.x) < n_prefill) {
    podp::prefill_body<T, 128>(pQ, pK, pV, pp, pi_qk, pi_pv, blockIdx.x);
  } else {
    podd::decode_body<16, NV, 128, T, 2>(dK, dV, dp, di_qk, di_pv, int(blockIdx.x) - n_prefill);
  }
}

}  // namespace

namespace fib200 {
template <int NV, typename T>
int pod_launch(const CUtensorMap& dK, const CUtensorMap& dV, const void* dparams, size_t dparams_size, uint32_t di_qk,
               uint32_t di_pv, int dgrid, int dsmem, bool pdl, cudaStream_t stream) {
  PodStage& st = pod_stage();
  st.have_prefill = false;  // consumed
  const bool f16 = std::is_same<T, __half>::value;
  FIB_CHECK(st.dqk == 128, "pod: the fused kernel is specialised for head_dim 128 on both sides");
  FIB_CHECK((st.f16 != 0) == f16, "pod: prefill and decode must use the same 16-bit dtype");
  FIB_CHECK(dparams_size == sizeof(podd::DecodeParams), "pod: decode parameter block mismatch");
  podp::PrefillParams pp;
  memcpy(&pp, st.params, sizeof(pp));
  podd::DecodeParams dp;
  memcpy(&dp, dparams, sizeof(dp));
  auto kern = pod_kernel<T, NV>;
  const int smem = st.smem > dsmem ? st.smem : dsmem;
  static bool attr_set = false;
  if (!attr_set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  LaunchCfg lc(dim3((unsigned)(st.grid + dgrid)), dim3(384), smem, stream, pdl);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, st.tm[0], st.tm[1], st.tm[2], pp, st.idesc[0], st.idesc[1], dK, dV, dp, di_qk,
                                    di_pv, st.grid));
  return 0;
}
}  // namespace fib200

// Arm / disarm the stage for the calling thread.  Armed: the next prefill_run parks, the next decode_paged_run fuses.
extern "C" int pod_arm(int64_t on) {
  pod_stage().armed = on != 0;
  pod_stage().have_prefill = false;
  return 0;
}

// out[0] = 1 when a prefill launch is parked (the prefill side took the tcgen05 path and can be fused).
extern "C" int pod_query(int64_t* out) {
  out[0] = pod_stage().have_prefill ? 1 : 0;
  return 0;
}
