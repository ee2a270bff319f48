// bootstrap.cuh — resampled means for simulation.summary() (reference: bootstrap_mean_ci, Stryke/stryke.py:143-162).
//
// The reference draws n_bootstrap resamples of len(data) values with np.random.choice and keeps their means; summary()
// calls it a few hundred times on series with one value per (iteration, day) — 1e11 gathered values for a C3-sized
// project, minutes of host time.  Here one block computes one resample: indices come from Philox4x32-10 (counter =
// (group of 4 indices, resample, 0, tag), key = seed) mapped to [0, n) by multiply-shift, values are gathered through
// L1/L2 (the series is at most a few MB) and summed in fp64.  Multiply-shift over-weights an index by at most n / 2^32
// (8.5e-6 for n = 36 500), far below the Monte Carlo error of the interval; the reference's own intervals are unseeded.
// Included at the end of passage.cu (uses its DevBuf / CK / fail helpers).
#pragma once

namespace stryke {

constexpr int BOOT_T = 256;

__global__ void __launch_bounds__(BOOT_T) bootstrap_means_kernel(const double* __restrict__ data, const uint32_t n, const int n_boot,
                                                                 const uint32_t k0, const uint32_t k1, double* __restrict__ means) {
  __shared__ double s_part[BOOT_T / 32];
  const uint32_t groups = (n + 3u) >> 2;
  for (int b = blockIdx.x; b < n_boot; b += gridDim.x) {
    double acc = 0.0;
    for (uint32_t g = threadIdx.x; g < groups; g += BOOT_T) {
      uint32_t w[4];
      Philox::block(g, (uint32_t)b, 0u, 0xB0075u, k0, k1, w);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4u * g + j < n) acc += __ldg(data + __umulhi(w[j], n));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < BOOT_T / 32; ++i) t += s_part[i];
      means[b] = t / (double)n;
    }
    __syncthreads();
  }
}

}  // namespace stryke

extern "C" int stryke_b200_bootstrap_means(const double* data, int64_t n, int32_t n_boot, uint64_t seed, int32_t device, double* means) {
  using namespace stryke;
  if (!data || !means || n <= 0 || n_boot <= 0 || n > 0xffffffffll) return fail(STRYKE_E_ARG, "bootstrap: bad argument");
  if (stryke_b200_device_count() <= 0) return fail(STRYKE_E_NOGPU, "no CUDA device visible: stryke_b200 has no CPU fallback");
  CK(cudaSetDevice(device));
  DevBuf d_data, d_means;
  CK(upload(d_data, data, (size_t)n));
  CK(d_means.alloc((size_t)n_boot * sizeof(double)));
  int n_sm = 0;
  if (device_sm_count(device, &n_sm)) return fail(STRYKE_E_CUDA, "cannot query the SM count");
  const int grid = std::min<int>(n_boot, n_sm * 8);
  bootstrap_means_kernel<<<grid, BOOT_T>>>(d_data.as<double>(), (uint32_t)n, n_boot, (uint32_t)seed, (uint32_t)(seed >> 32), d_means.as<double>());
  g_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpy(means, d_means.p, (size_t)n_boot * sizeof(double), cudaMemcpyDeviceToHost));
  return STRYKE_OK;
}
