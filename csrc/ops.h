// C++ entry points of the non-GEMM sm_100a kernels (elementwise.cu,
// optim_comm.cu, feed.cu).  All functions enqueue on the given stream and
// return the launch status; none of them synchronises.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfos {

// ---- elementwise.cu
cudaError_t bn_stats(const void* x, long long P, int C, float* sum, float* sumsq, cudaStream_t s);
cudaError_t bn_finalize(float* sum, float* sumsq, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float* mean, float* invstd,
                        float* scale, float* shift, int C, float count, float eps, float momentum,
                        cudaStream_t s);
cudaError_t bn_inference_coeffs(const float* gamma, const float* beta, const float* rm,
                                const float* rv, float* scale, float* shift, int C, float eps,
                                cudaStream_t s);
// mask (optional): one bit per element, set where the pre-activation value is > 0
// row order of the following bn_apply / bn_apply_finalize / bn_bwd_* launches: 1 = last row first
void bn_set_row_reverse(int flag);
cudaError_t bn_apply(const void* x, const void* residual, const float* scale, const float* shift,
                     void* y, uint8_t* mask, long long P, int C, int act, cudaStream_t s);
// Same, with the statistics finalisation folded in: every thread derives scale / shift of its
// channel group from the accumulated (sum, sumsq); one designated thread per group also stores
// mean / invstd / scale / shift (the backward kernels read them) and updates the running
// statistics.  The sums are NOT cleared here (other blocks are still reading them): the trainer
// zeroes its whole statistics arena once per step.
struct BnFinalize {
  const float* sum;
  const float* sumsq;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* mean;
  float* invstd;
  float* scale;
  float* shift;
  float count, eps, momentum;
};
cudaError_t bn_apply_finalize(const void* x, const void* residual, void* y, uint8_t* mask,
                              long long P, int C, int act, const BnFinalize& fin, cudaStream_t s);
cudaError_t bn_bwd_reduce(const void* dy, const void* x, const void* y, const float* mean,
                          const float* invstd, const float* fscale, const float* fshift,
                          long long P, int C, int relu, float* dgamma, float* dbeta,
                          cudaStream_t s);
cudaError_t bn_bwd_apply(const void* dy, const void* x, const void* y, const float* gamma,
                         const float* mean, const float* invstd, float* dgamma,
                         float* dbeta, const float* fscale, const float* fshift, void* dx,
                         void* dres, long long P, int C, int relu, const float* sum_g,
                         const float* sum_gx, cudaStream_t s);
cudaError_t add_act(const void* a, const void* b, void* out, long long n, int act, cudaStream_t s);
cudaError_t relu_bwd(const void* dy, const void* y, void* dx, long long n, cudaStream_t s);
cudaError_t colsum(const void* x, long long P, int C, float* out, cudaStream_t s);
cudaError_t maxpool_fwd(const void* x, void* y, uint8_t* idx, int N, int H, int W, int C, int OH,
                        int OW, int k, int stride, int pad, cudaStream_t s);
cudaError_t maxpool_bwd(const void* dy, const uint8_t* idx, void* dx, int N, int H, int W, int C,
                        int OH, int OW, int k, int stride, int pad, cudaStream_t s);
cudaError_t avgpool_fwd(const void* x, void* y, int N, int HW, int C, cudaStream_t s);
cudaError_t avgpool_bwd(const void* dy, void* dx, int N, int HW, int C, cudaStream_t s);
cudaError_t softmax_xent(const void* logits, int logits_fp32, const int* labels, void* dlogits,
                         float* loss_sum, float* correct_sum, long long rows, int V, int ld,
                         int ldd, float scale, cudaStream_t s);
cudaError_t decode_normalize(const uint8_t* in, void* out, int N, int H, int W, int C, int Wp,
                             int Cp, int wofs, const float* mean3, const float* istd3,
                             cudaStream_t s);
cudaError_t cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t s);
cudaError_t copy_channels(const void* src, void* dst, long long P, int C, int src_ld, int src_off,
                          int dst_ld, int dst_off, cudaStream_t s);
cudaError_t pixel_xent(const void* logits, const int* labels, void* dlogits, float* loss_sum,
                       float* correct_sum, long long P, int V, float scale, cudaStream_t s);

// ---- optim_comm.cu
constexpr int kMaxRanks = 8;
constexpr int kOptSgd = 0, kOptMomentum = 1, kOptAdam = 2;

struct AllreduceOptArgs {
  float* master;       // fp32 master weights (full length; a rank only touches its shards)
  float* state1;       // momentum / Adam m
  float* state2;       // Adam v
  const float* hyper;  // device: lr, momentum, wd, grad_scale, beta1, beta2, eps, step
  long long begin, end;  // element range of this bucket (begin % 8 == 0)
  long long decay_end;   // elements [0, decay_end) receive weight decay
  long long state_offset;
  int world, rank, slot, zero_grads;
  float* grads[kMaxRanks];             // every rank's fp32 gradient buffer (peer-mapped)
  __nv_bfloat16* weights[kMaxRanks];   // every rank's bf16 weight buffer (peer-mapped)
  uint32_t* flags[kMaxRanks];          // every rank's flag pad (peer-mapped)
  float* aux32[kMaxRanks];             // every rank's fp32 replica of elements >= aux_begin
  long long aux_begin;                 //   (BN scale/offset, biases); may be null
  const float* grads_mc;               // NVLS multicast views (optional)
  __nv_bfloat16* weights_mc;
  float* aux32_mc;
  uint32_t* epoch;          // local, one word per slot
  uint32_t* block_counter;  // local, one word per slot
};
// phase: 0 = fused; 1 = intra-host reduce-scatter only; 2 = update + all-gather only (optim_comm.cu)
cudaError_t allreduce_opt(const AllreduceOptArgs& a, int opt, int grid, cudaStream_t s, int phase = 0);

struct BcastArgs {
  int world, rank, root, slot;
  long long bytes;
  void* bufs[kMaxRanks];
  uint32_t* flags[kMaxRanks];
  uint32_t* epoch;
  uint32_t* block_counter;
};
cudaError_t bcast_pull(const BcastArgs& a, int grid, cudaStream_t s);
cudaError_t set_flag_timeout_ns(unsigned long long ns);  // bound of every device-side peer wait
cudaError_t flag_barrier(const BcastArgs& a, cudaStream_t s);
cudaError_t ps_push_dense(float* w_ps, const float* g, long long n, const float* hyper,
                          cudaStream_t s);
cudaError_t ps_push_sparse(float* w_ps, const float* g_rows, const int* idx, int nrows, int width,
                           const float* hyper, cudaStream_t s);
cudaError_t ps_pull(const float* w_ps, float* w_local, void* w_bf16, long long n, cudaStream_t s);

// fused tail of the ResNet stem (stem_fused.cu): BN + ReLU + 3x3/2 max pool forward; the max-pool
// backward folded into the BN backward reduction + apply (two launches)
cudaError_t stem_bn_relu_pool_fwd(const void* x, void* y, uint8_t* idx, int N, int H, int W, int C,
                                  int OH, int OW, const BnFinalize& fin, cudaStream_t s);
cudaError_t stem_pool_bn_bwd(const void* dy_pool, const uint8_t* idx, const void* x,
                             const float* gamma, const float* mean, const float* invstd,
                             const float* fscale, const float* fshift, float* dgamma, float* dbeta,
                             void* dx, int N, int H, int W, int C, int OH, int OW, cudaStream_t s);

// parameter server with resident optimizer state ("slot mode", parallel/ps.py)
struct PsApplyArgs {          // server side, all pointers local to the PS GPU
  float* master;              // [n] fp32 parameters of this server's slice (+ non-trainable tail)
  float* state1;              // momentum / Adam m (may be null for SGD)
  float* state2;              // Adam v
  __nv_bfloat16* wbf16;       // [n] bf16 serving copy that workers pull
  const float* slot;          // [n] one worker's gradient slot
  const float* hyper;         // lr, momentum, wd, grad_scale, beta1, beta2, eps, step
  long long n, lo;            // slice length and its global offset
  long long decay_end;        // global: elements below receive weight decay
  long long ema_begin;        // global: elements from here on are running statistics
  uint32_t* applied_flag;     // this slot's "applied" word
  uint32_t seq;
  uint32_t* block_counter;
};
struct PsPushArgs {           // worker side
  float* slot;                // peer-mapped: this worker's slot on the PS GPU
  const float* grads;         // local fp32 gradients [total]
  const float* running;       // local running statistics after the step [n_running]
  const float* running_pulled;  // their values as pulled before the step
  long long n, lo, total;     // slice length / global offset; total = number of parameters
  const uint32_t* applied_flag;  // peer-mapped
  uint32_t need_applied;
  uint32_t* ready_flag;       // peer-mapped
  uint32_t seq;
  uint32_t* block_counter;    // local
};
struct PsPullArgs {           // worker side
  const __nv_bfloat16* wbf16;   // peer-mapped bf16 serving copy of the slice
  const float* master;        // peer-mapped fp32 slice
  __nv_bfloat16* weights;     // local bf16 weights [total]
  float* aux32;               // local fp32 replica of [decay_end, total)
  float* running;             // local running statistics
  float* running_pulled;      // snapshot for the next push
  long long n, lo, total, decay_end;
};
cudaError_t ps_apply(const PsApplyArgs& a, int opt, cudaStream_t s);
cudaError_t ps_push_slot(const PsPushArgs& a, int grid, cudaStream_t s);
cudaError_t ps_pull_model(const PsPullArgs& a, int grid, cudaStream_t s);

// ---- small direct kernels (smallops.cu): MNIST conv (Cin=1), depthwise 3x3
cudaError_t conv3x3_c1_fwd(const void* x, const void* w, const float* bias, void* y, int N, int H,
                           int W, int Cout, int relu, cudaStream_t s);
cudaError_t conv3x3_c1_wgrad(const void* x, const void* dy, float* dw, float* dbias, int N, int H,
                             int W, int Cout, cudaStream_t s);
cudaError_t depthwise3x3_fwd(const void* x, const void* w, const float* bias, void* y, int N,
                             int H, int W, int C, int stride, int act, cudaStream_t s);

}  // namespace tfos
