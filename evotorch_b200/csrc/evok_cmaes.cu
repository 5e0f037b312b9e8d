// CMA-ES generation glue (cmaes.py:432-553 of the reference): the D-vector / N-vector arithmetic between the dense
// contractions, fused so that a whole generation is a short chain of kernels without host reads and is CUDA-graph
// capturable.  The reference runs these as ~35 eager torch ops per generation.
//
//   evok_cmaes_row_weights   N-vector: positive part of the assigned weights (recombination) and the active-CMA
//                            reweighting  w_i > 0 ? w_i : d * w_i / ||z_i||^2   (cmaes.py:468-475, :531-535); one pass over Z
//   evok_cmaes_vector_update D-vectors + scalars, one CTA: m, p_sigma, sigma, h_sig, p_c and the three coefficients of the
//                            covariance update consumed by evok_gemm_nt_affine (cmaes.py:454-517, :31-46, :537-545)
#include "evok_common.cuh"

namespace evok {

// one warp per row: ||z_i||^2, then the two weight vectors
__global__ void __launch_bounds__(256) cmaes_row_weights_kernel(const float* __restrict__ aw, const float* __restrict__ Z, int64_t ldz, int64_t N,
                                                                int64_t D, int active, float* __restrict__ w_pos, float* __restrict__ w_act) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const float a = aw[row];
  float out_act = a;
  if (active && !(a > 0.0f)) {  // only the non-positive weights need the row norm (cmaes.py:532)
    const float* z = Z + row * ldz;
    float s = 0.0f;
    if ((D & 3) == 0 && (ldz & 3) == 0 && aligned16_dev(Z)) {
      for (int64_t q = lane; q < (D >> 2); q += 32) {
        const float4 v = ld_stream4(z + 4 * q);
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
      }
    } else {
      for (int64_t j = lane; j < D; j += 32) {
        const float v = z[j];
        s = fmaf(v, v, s);
      }
    }
    s = warp_sum(s);
    out_act = __fdiv_rn((float)D * a, s);
  }
  if (lane == 0) {
    w_pos[row] = fmaxf(a, 0.0f);
    w_act[row] = out_act;
  }
}

struct CmaesConsts {
  float c_m, c_sigma, damp_sigma, c_c, c_1, c_mu, vd_sigma, vd_c, unbiased_expectation, weights_sum;
  int csa_squared;
};

constexpr int kCmaThreads = 1024;

__global__ void __launch_bounds__(kCmaThreads)
    cmaes_vector_update_kernel(const float* __restrict__ local_disp, const float* __restrict__ shaped_disp, int64_t D, float* __restrict__ m,
                               float* __restrict__ p_sigma, float* __restrict__ p_c, float* __restrict__ sigma, long long* steps_dev,
                               long long steps_host, const __grid_constant__ CmaesConsts c, float* __restrict__ k_out, float* __restrict__ h_sig_out) {
  __shared__ double sm[33];
  const float sig = *sigma;
  const long long steps = steps_dev ? *steps_dev : steps_host;
  // update_m (cmaes.py:477-479, uses the OLD sigma) and update_p_sigma (:483-490)
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < D; i += kCmaThreads) {
    m[i] = m[i] + c.c_m * sig * shaped_disp[i];
    const float ps = (1.0f - c.c_sigma) * p_sigma[i] + c.vd_sigma * local_disp[i];
    p_sigma[i] = ps;
    acc += (double)ps * (double)ps;
  }
  const float pnorm = (float)sqrt(block_sum<double>(acc, sm));
  // update_sigma (:492-507)
  const float dn = (float)D;
  const float expo = c.csa_squared ? (pnorm * pnorm / dn - 1.0f) * 0.5f : pnorm / c.unbiased_expectation - 1.0f;
  const float new_sigma = sig * expf((c.c_sigma / c.damp_sigma) * expo);
  // _h_sig (:31-46): generation counter BEFORE the increment
  const double decay = 1.0 - pow(1.0 - (double)c.c_sigma, (double)(2 * steps + 1));
  const float squared_sum = (float)((double)(pnorm * pnorm) / decay);
  const float h = ((squared_sum / dn) - 1.0f < 1.0f + 4.0f / (dn + 1.0f)) ? 1.0f : 0.0f;
  // update_p_c (:509-517)
  for (int64_t i = threadIdx.x; i < D; i += kCmaThreads) p_c[i] = (1.0f - c.c_c) * p_c[i] + h * c.vd_c * shaped_disp[i];
  if (threadIdx.x == 0) {
    *sigma = new_sigma;
    if (steps_dev) *steps_dev = steps + 1;
    // covariance update C <- C + c1a (pc pc^T - C) + c_mu (S - sum(w) C), pc = weighted_pc * p_c   (:537-549)
    const float c1a = c.c_1 * (1.0f - (1.0f - h * h) * c.c_c * (2.0f - c.c_c));
    const float wpc2 = c.c_1 / (c1a + 1e-23f);  // weighted_pc squared
    k_out[0] = c.c_mu;
    k_out[1] = 1.0f - c1a - c.c_mu * c.weights_sum;
    k_out[2] = c1a * wpc2;
    if (h_sig_out) *h_sig_out = h;
  }
}

}  // namespace evok

using namespace evok;

extern "C" EVOK_API int evok_cmaes_row_weights(const float* assigned_weights, const float* Z, int64_t ldz, int64_t N, int64_t D, int active,
                                               float* w_positive, float* w_active, void* stream) {
  if (!assigned_weights || !Z || !w_positive || !w_active) return EVOK_E_NULLPTR;
  if (N <= 0 || D <= 0 || ldz < D) return EVOK_E_BADSIZE;
  cmaes_row_weights_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(assigned_weights, Z, ldz, N, D, active, w_positive, w_active);
  EVOK_CHECK_LAUNCH();
  return 0;
}

extern "C" EVOK_API int evok_cmaes_vector_update(const float* local_disp, const float* shaped_disp, int64_t D, float* m, float* p_sigma, float* p_c,
                                                 float* sigma_dev, int64_t* steps_dev, int64_t steps_host, const float* consts_host, int csa_squared,
                                                 float* k_out, float* h_sig_out, void* stream) {
  if (!local_disp || !shaped_disp || !m || !p_sigma || !p_c || !sigma_dev || !consts_host || !k_out) return EVOK_E_NULLPTR;
  if (D <= 0) return EVOK_E_BADSIZE;
  CmaesConsts c;
  c.c_m = consts_host[0]; c.c_sigma = consts_host[1]; c.damp_sigma = consts_host[2]; c.c_c = consts_host[3]; c.c_1 = consts_host[4];
  c.c_mu = consts_host[5]; c.vd_sigma = consts_host[6]; c.vd_c = consts_host[7]; c.unbiased_expectation = consts_host[8];
  c.weights_sum = consts_host[9];
  c.csa_squared = csa_squared;
  cmaes_vector_update_kernel<<<1, kCmaThreads, 0, (cudaStream_t)stream>>>(local_disp, shaped_disp, D, m, p_sigma, p_c, sigma_dev,
                                                                         reinterpret_cast<long long*>(steps_dev), (long long)steps_host, c, k_out,
                                                                         h_sig_out);
  EVOK_CHECK_LAUNCH();
  return 0;
}
