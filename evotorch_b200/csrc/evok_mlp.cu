// K8: batched flat-parameter MLP policy forward -- one observation per policy (the reference's `Policy.__call__`,
// neuroevolution/net/vecrl.py:1240-1279, which vmaps `functional_call` over the rows of an N x L parameter matrix).
// Every policy has its OWN weights, so this is a batched GEMV: 0.5 flop per parameter byte, i.e. purely HBM-read bound
// (26.4 GB of parameters at N = 65 536, L = 100 881).  Tensor cores cannot help: no operand is shared between rows.
//
// Layout of a parameter row (net/functional.py:118-129, torch.nn.Linear order): for each layer, W (out x in, row-major)
// then b (out).  L is odd in general (100 881), so rows are only 4-byte aligned: the kernel reads weights with coalesced
// 32-bit loads (a warp covers 128 contiguous bytes per instruction, 4 neuron rows in flight per warp).
// One CTA per policy (persistent grid-stride); activations ping-pong through shared memory.
#include "evok_common.cuh"

namespace evok {

constexpr int kMlpThreads = 256;
constexpr int kMlpWarps = kMlpThreads / 32;
constexpr int kMlpMaxLayers = 8;
constexpr int kMlpMaxWidth = 2048;
constexpr int kMlpNeuronsPerPass = 4;

struct MlpSpec {
  int n_layers;
  int dims[kMlpMaxLayers + 1];
  int acts[kMlpMaxLayers];
  int64_t w_off[kMlpMaxLayers];  // offset of W_l inside a parameter row; b_l follows at w_off + in*out
  int max_width;
};

__device__ __forceinline__ float activate(float v, int act) {
  switch (act) {
    case EVOK_ACT_TANH: return tanhf(v);
    case EVOK_ACT_RELU: return fmaxf(v, 0.0f);
    case EVOK_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

__global__ void __launch_bounds__(kMlpThreads)
    mlp_forward_kernel(const float* __restrict__ params, int64_t ldp, const float* __restrict__ obs, int64_t ldo, float* __restrict__ out,
                       int64_t ldout, int64_t N, const __grid_constant__ MlpSpec spec) {
  extern __shared__ float act_buf[];  // 2 x max_width
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
    const float* prow = params + i * ldp;
    float* cur = act_buf;
    float* nxt = act_buf + spec.max_width;
    for (int k = threadIdx.x; k < spec.dims[0]; k += kMlpThreads) cur[k] = ld_stream1(obs + i * ldo + k);
    __syncthreads();
    for (int l = 0; l < spec.n_layers; ++l) {
      const int n_in = spec.dims[l], n_out = spec.dims[l + 1];
      const float* W = prow + spec.w_off[l];
      const float* b = W + (int64_t)n_in * n_out;
      const bool last = l == spec.n_layers - 1;
      for (int j0 = warp * kMlpNeuronsPerPass; j0 < n_out; j0 += kMlpWarps * kMlpNeuronsPerPass) {
        float acc[kMlpNeuronsPerPass];
#pragma unroll
        for (int t = 0; t < kMlpNeuronsPerPass; ++t) acc[t] = 0.0f;
        const int n_here = min(kMlpNeuronsPerPass, n_out - j0);
        if (n_here == kMlpNeuronsPerPass) {
          const float* w0 = W + (int64_t)j0 * n_in;
          for (int k = lane; k < n_in; k += 32) {
            const float x = cur[k];
#pragma unroll
            for (int t = 0; t < kMlpNeuronsPerPass; ++t) acc[t] = fmaf(ld_stream1(w0 + (int64_t)t * n_in + k), x, acc[t]);
          }
        } else {
          for (int t = 0; t < n_here; ++t) {
            const float* w0 = W + (int64_t)(j0 + t) * n_in;
            for (int k = lane; k < n_in; k += 32) acc[t] = fmaf(ld_stream1(w0 + k), cur[k], acc[t]);
          }
        }
#pragma unroll
        for (int t = 0; t < kMlpNeuronsPerPass; ++t) acc[t] = warp_sum(acc[t]);
        if (lane < n_here) {
          float v = acc[0];
#pragma unroll
          for (int t = 1; t < kMlpNeuronsPerPass; ++t) v = lane == t ? acc[t] : v;
          v = activate(v + ld_stream1(b + j0 + lane), spec.acts[l]);
          if (last) out[i * ldout + j0 + lane] = v;
          else nxt[j0 + lane] = v;
        }
      }
      __syncthreads();
      float* tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
  }
}

}  // namespace evok

using namespace evok;

extern "C" EVOK_API int64_t evok_mlp_parameter_length(int n_layers, const int32_t* dims_host) {
  if (!dims_host || n_layers < 1 || n_layers > kMlpMaxLayers) return -1;
  int64_t total = 0;
  for (int l = 0; l < n_layers; ++l) total += (int64_t)dims_host[l] * dims_host[l + 1] + dims_host[l + 1];
  return total;
}

extern "C" EVOK_API int evok_mlp_forward(const float* params, int64_t ldp, const float* obs, int64_t ldo, float* out, int64_t ldout,
                                         int64_t N, int n_layers, const int32_t* dims_host, const int32_t* acts_host, void* stream) {
  if (!params || !obs || !out || !dims_host || !acts_host) return EVOK_E_NULLPTR;
  if (n_layers < 1 || n_layers > kMlpMaxLayers || N < 0) return EVOK_E_BADSIZE;
  MlpSpec spec;
  spec.n_layers = n_layers;
  int64_t off = 0;
  int maxw = 0;
  for (int l = 0; l <= n_layers; ++l) {
    const int d = dims_host[l];
    if (d < 1 || d > kMlpMaxWidth) return EVOK_E_BADSIZE;
    spec.dims[l] = d;
    if (d > maxw) maxw = d;
  }
  for (int l = 0; l < n_layers; ++l) {
    if (acts_host[l] < EVOK_ACT_NONE || acts_host[l] > EVOK_ACT_SIGMOID) return EVOK_E_BADENUM;
    spec.acts[l] = acts_host[l];
    spec.w_off[l] = off;
    off += (int64_t)spec.dims[l] * spec.dims[l + 1] + spec.dims[l + 1];
  }
  spec.max_width = maxw;
  if (ldp < off || ldo < spec.dims[0] || ldout < spec.dims[n_layers]) return EVOK_E_BADSIZE;
  if (N == 0) return 0;
  const size_t smem = 2 * (size_t)maxw * sizeof(float);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mlp_forward_kernel, kMlpThreads, smem) != cudaSuccess || per_sm <= 0) per_sm = 4;
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (int64_t)per_sm * sms;
  if (grid > N) grid = N;
  mlp_forward_kernel<<<(unsigned)grid, kMlpThreads, smem, (cudaStream_t)stream>>>(params, ldp, obs, ldo, out, ldout, N, spec);
  EVOK_CHECK_LAUNCH();
  return 0;
}
