// internal.h — host-side glue between the C ABI (api.cu) and the per-size kernel instantiations.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace b2l {

enum KernelOp : int { OP_LAUNCH = 0, OP_SET_SMEM = 1, OP_OCCUPANCY = 2 };

// One translation unit per LOG2M (fwd_inst.cu / inv_inst.cu compiled with -DB2L_LOG2M=k) exports these.
#define B2L_DECL_FWD(L) \
  cudaError_t fwd_op_##L(int op, int nw, int mode, const FwdArgs* a, int grid, size_t smem, cudaStream_t st, int* result);
#define B2L_DECL_INV(L) \
  cudaError_t inv_op_##L(int op, int nw, const InvArgs* a, int grid, size_t smem, cudaStream_t st, int* result);

B2L_DECL_FWD(2) B2L_DECL_FWD(3) B2L_DECL_FWD(4) B2L_DECL_FWD(5) B2L_DECL_FWD(6) B2L_DECL_FWD(7)
B2L_DECL_FWD(8) B2L_DECL_FWD(9) B2L_DECL_FWD(10) B2L_DECL_FWD(11) B2L_DECL_FWD(12)
B2L_DECL_INV(2) B2L_DECL_INV(3) B2L_DECL_INV(4) B2L_DECL_INV(5) B2L_DECL_INV(6) B2L_DECL_INV(7)
B2L_DECL_INV(8) B2L_DECL_INV(9) B2L_DECL_INV(10) B2L_DECL_INV(11) B2L_DECL_INV(12)

struct CztArgs;
struct CztInvArgs;
#define B2L_DECL_CZT(L)                                                                                    \
  cudaError_t czt_op_##L(int op, const CztArgs* a, int grid, size_t smem, cudaStream_t st, int* result); \
  cudaError_t czt_inv_op_##L(int op, const CztInvArgs* a, int grid, size_t smem, cudaStream_t st, int* result);
B2L_DECL_CZT(5) B2L_DECL_CZT(6) B2L_DECL_CZT(7) B2L_DECL_CZT(8) B2L_DECL_CZT(9) B2L_DECL_CZT(10) B2L_DECL_CZT(11) B2L_DECL_CZT(12)

constexpr int kMinLog2M = 2, kMaxLog2M = 12;   // n_fft = 2^(LOG2M+1): 8 .. 8192

// Host mirror of FftCfg<LOG2M, TPF> (fft_engine.cuh): same schedule, evaluated at run time.
struct HostFftCfg {
  int log2m, M, tpf, ppt, logp, npass;
  explicit HostFftCfg(int l2m) {
    log2m = l2m;
    M = 1 << l2m;
    tpf = M >= 32 ? M / 32 : 1;
    ppt = M / tpf;
    logp = 0;
    while ((1 << logp) < ppt) ++logp;
    npass = ppt == 1 ? 1 : (log2m + logp - 1) / logp;
  }
  int log_radix(int s) const { int left = log2m - s * logp; return left >= logp ? logp : left; }
  int radix(int s) const { return 1 << log_radix(s); }
  int sublen(int s) const { return 1 << (s * logp); }
  int tw_offset(int s) const {
    int off = 0;
    for (int q = 1; q < s; ++q) off += (radix(q) - 1) * sublen(q);
    return off;
  }
  int tw_count() const { return tw_offset(npass); }
  int xbuf_f2() const { return M + M / 32; }
  // warps per CTA of czt_kernel for P = 2^log2p (mirror of czt_inst.cu)
  int czt_nw() const { return log2m >= 10 ? 16 : (tpf > 16 ? 16 : tpf); }
  // warps per CTA tried in order (first that fits shared memory wins)
  int nw_options(int out[2]) const {
    if (log2m >= 10) { out[0] = 16; out[1] = 8; return 2; }
    out[0] = tpf >= 1 ? (tpf > 16 ? 16 : tpf) : 1;
    return 1;
  }
};

}  // namespace b2l

// ---- glue for translation units other than api.cu (f64_api.cu): the context stays opaque to them
struct b2l_ctx;
cudaStream_t b2l_internal_stream(b2l_ctx* c);
int b2l_internal_device(b2l_ctx* c);
int* b2l_internal_status(b2l_ctx* c);
size_t b2l_internal_smem_optin(b2l_ctx* c);
int b2l_internal_sm_count(b2l_ctx* c);
void b2l_internal_count_launches(b2l_ctx* c, int n);
int b2l_internal_fail(int code, const char* fmt, ...);
