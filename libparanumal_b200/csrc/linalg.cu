// linAlg_t on the device: the 11 streaming vector operations and the reductions of
// include/linAlg.hpp:52-120 (kernels libs/linAlg/okl/*.okl; host side libs/linAlg/linAlg.cpp:37-224).
//
// All kernels are pure HBM streams: 128-bit loads/stores where the pointers allow, grid sized
// to a multiple of the SM count, grid-stride loops.  Reductions are two-level with a FIXED
// block count and a fixed tree inside each block, so results are run-to-run deterministic; the
// cross-rank step is a device-side NCCL all-reduce of the scalar (no pinned-host MPI staging).
// beta == 0 variants never read y (callers pass uninitialised outputs).
#include <memory>
#include <unordered_map>

#include "linalg.hpp"

using namespace libp_b200;

namespace {

constexpr int kBlock = 256;

inline int vec_grid(dlong N, int per_thread) {
  long b = ((long)N + (long)kBlock * per_thread - 1) / ((long)kBlock * per_thread);
  const long cap = (long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Generic element-wise driver: F(n) applied with double2 vectorisation on the aligned body.
// OP::apply2 works on pairs, OP::apply1 on the tail.
template <class OP>
__global__ void __launch_bounds__(kBlock) ew_kernel(dlong N, OP op) {
  const dlong N2 = N >> 1;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N2; n += gridDim.x * kBlock) op.pair(n);
  if (blockIdx.x == 0 && threadIdx.x == 0 && (N & 1)) op.one(N - 1);
}
template <class OP>
__global__ void __launch_bounds__(kBlock) ew_scalar_kernel(dlong N, OP op) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) op.one(n);
}
template <class OP>
void run_ew(dlong N, const OP& op, bool vec_ok, cudaStream_t s) {
  if (N <= 0) return;
  if (vec_ok) ew_kernel<OP><<<vec_grid(N, 2), kBlock, 0, s>>>(N, op);
  else ew_scalar_kernel<OP><<<vec_grid(N, 1), kBlock, 0, s>>>(N, op);
  CUDA_CHECK(cudaGetLastError());
}

#define D2(p) reinterpret_cast<double2*>(p)
#define CD2(p) reinterpret_cast<const double2*>(p)

struct SetOp { double alpha; double* a;
  __device__ void pair(dlong n) const { D2(a)[n] = make_double2(alpha, alpha); }
  __device__ void one(dlong n) const { a[n] = alpha; } };
struct AddOp { double alpha; double* a;
  __device__ void pair(dlong n) const { double2 v = D2(a)[n]; v.x += alpha; v.y += alpha; D2(a)[n] = v; }
  __device__ void one(dlong n) const { a[n] += alpha; } };
struct ScaleOp { double alpha; double* a;
  __device__ void pair(dlong n) const { double2 v = D2(a)[n]; v.x *= alpha; v.y *= alpha; D2(a)[n] = v; }
  __device__ void one(dlong n) const { a[n] *= alpha; } };
// z = alpha*x + beta*y  (y may alias z); kReadY=false never touches y
template <bool kReadY>
struct AxpyOp { double alpha; const double* x; double beta; const double* y; double* z;
  __device__ void pair(dlong n) const {
    const double2 xv = CD2(x)[n]; double2 r;
    if (kReadY) { const double2 yv = CD2(y)[n]; r.x = alpha * xv.x + beta * yv.x; r.y = alpha * xv.y + beta * yv.y; }
    else { r.x = alpha * xv.x; r.y = alpha * xv.y; }
    D2(z)[n] = r; }
  __device__ void one(dlong n) const { z[n] = kReadY ? alpha * x[n] + beta * y[n] : alpha * x[n]; } };
// z = alpha*a*x + beta*y  (or alpha*x/a when kDiv)
template <bool kReadY, bool kDiv>
struct AmxpyOp { double alpha; const double* a; const double* x; double beta; const double* y; double* z;
  __device__ double f(double av, double xv) const { return kDiv ? alpha * xv / av : alpha * av * xv; }
  __device__ void pair(dlong n) const {
    const double2 av = CD2(a)[n], xv = CD2(x)[n]; double2 r;
    r.x = f(av.x, xv.x); r.y = f(av.y, xv.y);
    if (kReadY) { const double2 yv = CD2(y)[n]; r.x += beta * yv.x; r.y += beta * yv.y; }
    D2(z)[n] = r; }
  __device__ void one(dlong n) const { double r = f(a[n], x[n]); if (kReadY) r += beta * y[n]; z[n] = r; } };

// ---------------------------------------------------------------- reductions
enum RedKind { kSum = 0, kMin = 1, kMax = 2 };

template <int kind> __device__ inline double red_id();
template <> __device__ inline double red_id<kSum>() { return 0.0; }
template <> __device__ inline double red_id<kMin>() { return 1.7976931348623157e+308; }
template <> __device__ inline double red_id<kMax>() { return -1.7976931348623157e+308; }
template <int kind> __device__ inline double red_op(double a, double b) {
  if (kind == kSum) return a + b;
  if (kind == kMin) return b < a ? b : a;
  return b > a ? b : a;
}

template <int kind>
__device__ inline double block_reduce(double v) {
  __shared__ double s_w[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_op<kind>(v, __shfl_down_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_w[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? s_w[threadIdx.x] : red_id<kind>();
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op<kind>(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  return v;  // valid in thread 0
}

// mode: 0 a ; 1 x*y ; 2 w*x*y ; 3 a*a ; 4 w*a*a
template <int kind, int mode>
__global__ void __launch_bounds__(kRedBlock) red1_kernel(dlong N, const double* __restrict__ w,
                                                         const double* __restrict__ x, const double* __restrict__ y,
                                                         double* __restrict__ partials) {
  double acc = red_id<kind>();
  for (dlong n = blockIdx.x * kRedBlock + threadIdx.x; n < N; n += gridDim.x * kRedBlock) {
    double v;
    if (mode == 0) v = x[n];
    else if (mode == 1) v = x[n] * y[n];
    else if (mode == 2) v = w[n] * x[n] * y[n];
    else if (mode == 3) v = x[n] * x[n];
    else v = w[n] * x[n] * x[n];
    acc = red_op<kind>(acc, v);
  }
  acc = block_reduce<kind>(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
template <int kind>
__global__ void __launch_bounds__(kRedMaxBlocks) red2_kernel(int nparts, const double* __restrict__ partials,
                                                             double* __restrict__ out) {
  double v = (threadIdx.x < nparts) ? partials[threadIdx.x] : red_id<kind>();
  v = block_reduce<kind>(v);
  if (threadIdx.x == 0) out[0] = v;
}

template <int kind, int mode>
void reduce_dev(dlong N, const double* w, const double* x, const double* y, double* d_out, cudaStream_t s) {
  RedScratch& rs = red_scratch(s);
  rs.ensure();
  const int nb = red_blocks(N);
  red1_kernel<kind, mode><<<nb, kRedBlock, 0, s>>>(N, w, x, y, rs.partials.p);
  red2_kernel<kind><<<1, kRedMaxBlocks, 0, s>>>(nb, rs.partials.p, d_out);
  CUDA_CHECK(cudaGetLastError());
}

template <int kind, int mode>
double reduce_host(dlong N, const double* w, const double* x, const double* y, libp_comm_t comm, cudaStream_t s) {
  RedScratch& rs = red_scratch(s);
  rs.ensure();
  reduce_dev<kind, mode>(N, w, x, y, rs.result.p, s);
  if (comm && comm->size > 1) comm->allreduce_dev(rs.result.p, 1, kind == kSum ? LIBP_ADD : kind == kMin ? LIBP_MIN : LIBP_MAX, s);
  CUDA_CHECK(cudaMemcpyAsync(rs.h_result, rs.result.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  return rs.h_result[0];
}

}  // namespace

namespace libp_b200 {
void RedScratch::ensure() {
  if (!partials.p) {
    partials.alloc((size_t)kRedMaxBlocks * 4);
    result.alloc(8);
    CUDA_CHECK(cudaMallocHost(&h_result, 8 * sizeof(double)));
  }
}
RedScratch::~RedScratch() {
  if (h_result) cudaFreeHost(h_result);
}
// one scratch per (host thread, stream): reductions issued concurrently from different streams or threads never share
// partial sums or landing slots
RedScratch& red_scratch(cudaStream_t s) {
  thread_local std::unordered_map<cudaStream_t, std::unique_ptr<RedScratch>> table;
  std::unique_ptr<RedScratch>& p = table[s];
  if (!p) p.reset(new RedScratch());
  return *p;
}
void dot_to_device(dlong N, const double* x, const double* y, const double* w, double* d_out, cudaStream_t s) {
  if (w) reduce_dev<kSum, 2>(N, w, x, y, d_out, s);
  else if (y) reduce_dev<kSum, 1>(N, nullptr, x, y, d_out, s);
  else reduce_dev<kSum, 0>(N, nullptr, x, nullptr, d_out, s);
}
}  // namespace libp_b200

#define VEC_API(NAME, CHECKS, BODY)            \
  LIBP_API_BEGIN                               \
  LIBP_CHECK(N >= 0, "negative length");       \
  cudaStream_t s = as_stream(stream);          \
  (void)s;                                     \
  if (N > 0) { CHECKS; BODY; }                 \
  LIBP_API_END

extern "C" {

int libp_linalg_set(libp_dlong N, libp_dfloat alpha, libp_dfloat* a, void* stream) {
  VEC_API(set, LIBP_CHECK(a, "null pointer"), run_ew(N, SetOp{alpha, a}, aligned16(a), s))
}
int libp_linalg_add(libp_dlong N, libp_dfloat alpha, libp_dfloat* a, void* stream) {
  VEC_API(add, LIBP_CHECK(a, "null pointer"), run_ew(N, AddOp{alpha, a}, aligned16(a), s))
}
int libp_linalg_scale(libp_dlong N, libp_dfloat alpha, libp_dfloat* a, void* stream) {
  VEC_API(scale, LIBP_CHECK(a, "null pointer"), run_ew(N, ScaleOp{alpha, a}, aligned16(a), s))
}
int libp_linalg_axpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* x, libp_dfloat beta, libp_dfloat* y,
                     void* stream) {
  VEC_API(axpy, LIBP_CHECK(x && y, "null pointer"),
          if (beta != 0.0) run_ew(N, AxpyOp<true>{alpha, x, beta, y, y}, aligned16(x) && aligned16(y), s);
          else run_ew(N, AxpyOp<false>{alpha, x, beta, y, y}, aligned16(x) && aligned16(y), s))
}
int libp_linalg_zaxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* x, libp_dfloat beta, const libp_dfloat* y,
                      libp_dfloat* z, void* stream) {
  VEC_API(zaxpy, LIBP_CHECK(x && y && z, "null pointer"),
          run_ew(N, AxpyOp<true>{alpha, x, beta, y, z}, aligned16(x) && aligned16(y) && aligned16(z), s))
}
int libp_linalg_amx(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, libp_dfloat* x, void* stream) {
  VEC_API(amx, LIBP_CHECK(a && x, "null pointer"),
          run_ew(N, AmxpyOp<false, false>{alpha, a, x, 0.0, x, x}, aligned16(a) && aligned16(x), s))
}
int libp_linalg_amxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                      libp_dfloat* y, void* stream) {
  VEC_API(amxpy, LIBP_CHECK(a && x && y, "null pointer"),
          const bool v = aligned16(a) && aligned16(x) && aligned16(y);
          if (beta != 0.0) run_ew(N, AmxpyOp<true, false>{alpha, a, x, beta, y, y}, v, s);
          else run_ew(N, AmxpyOp<false, false>{alpha, a, x, beta, y, y}, v, s))
}
int libp_linalg_zamxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                       const libp_dfloat* y, libp_dfloat* z, void* stream) {
  VEC_API(zamxpy, LIBP_CHECK(a && x && y && z, "null pointer"),
          run_ew(N, AmxpyOp<true, false>{alpha, a, x, beta, y, z},
                 aligned16(a) && aligned16(x) && aligned16(y) && aligned16(z), s))
}
int libp_linalg_adx(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, libp_dfloat* x, void* stream) {
  VEC_API(adx, LIBP_CHECK(a && x, "null pointer"),
          run_ew(N, AmxpyOp<false, true>{alpha, a, x, 0.0, x, x}, aligned16(a) && aligned16(x), s))
}
int libp_linalg_adxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                      libp_dfloat* y, void* stream) {
  VEC_API(adxpy, LIBP_CHECK(a && x && y, "null pointer"),
          const bool v = aligned16(a) && aligned16(x) && aligned16(y);
          if (beta != 0.0) run_ew(N, AmxpyOp<true, true>{alpha, a, x, beta, y, y}, v, s);
          else run_ew(N, AmxpyOp<false, true>{alpha, a, x, beta, y, y}, v, s))
}
int libp_linalg_zadxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                       const libp_dfloat* y, libp_dfloat* z, void* stream) {
  VEC_API(zadxpy, LIBP_CHECK(a && x && y && z, "null pointer"),
          run_ew(N, AmxpyOp<true, true>{alpha, a, x, beta, y, z},
                 aligned16(a) && aligned16(x) && aligned16(y) && aligned16(z), s))
}

#define RED_API(KIND, MODE, W, X, Y, POST)                                            \
  LIBP_API_BEGIN                                                                       \
  LIBP_CHECK(result != nullptr, "null result");                                        \
  LIBP_CHECK(N >= 0, "negative length");                                               \
  double r = reduce_host<KIND, MODE>(N, W, X, Y, comm, as_stream(stream));             \
  *result = POST;                                                                      \
  LIBP_API_END

int libp_linalg_min(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result) {
  RED_API(kMin, 0, nullptr, a, nullptr, r)
}
int libp_linalg_max(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result) {
  RED_API(kMax, 0, nullptr, a, nullptr, r)
}
int libp_linalg_sum(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result) {
  RED_API(kSum, 0, nullptr, a, nullptr, r)
}
int libp_linalg_norm2(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result) {
  RED_API(kSum, 3, nullptr, a, nullptr, sqrt(r))
}
int libp_linalg_inner_prod(libp_dlong N, const libp_dfloat* x, const libp_dfloat* y, libp_comm_t comm, void* stream,
                           libp_dfloat* result) {
  RED_API(kSum, 1, nullptr, x, y, r)
}
int libp_linalg_weighted_norm2(libp_dlong N, const libp_dfloat* w, const libp_dfloat* a, libp_comm_t comm,
                               void* stream, libp_dfloat* result) {
  RED_API(kSum, 4, w, a, nullptr, sqrt(r))
}
int libp_linalg_weighted_inner_prod(libp_dlong N, const libp_dfloat* w, const libp_dfloat* x, const libp_dfloat* y,
                                    libp_comm_t comm, void* stream, libp_dfloat* result) {
  RED_API(kSum, 2, w, x, y, r)
}

}  // extern "C"
