/* synth_gen.c — TEST / BENCH INFRASTRUCTURE (oracle/): the corpus recipe of bench_data.py in C + OpenMP, so that
 * the CPU reference arm of bench.py can materialise its 10M x 384 fp32 rows on the host in seconds instead of
 * minutes.  Bit-identical to bench_data.rows_np / rows_torch (tests/test_bench_cpu.py checks it): the same
 * splitmix64 finaliser, Irwin-Hall sums of four 16-bit fields, one fp32 multiply and one fp32 add per element.
 * Not part of the product path; built by __graft_entry__.build() into oracle/_build/libfrs_synth.so. */
#include <stdint.h>
#include <stddef.h>

static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline float irwin_hall(uint64_t h) {
  const uint64_t s = (h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48);
  return (float)s - 131070.0f;
}

/* x[m][dim] = cent[cid(r)] + field(r, c) * noise_scale ; ticker[m] = searchsorted(cdf, u(r), right) */
void frs_synth_rows(uint64_t row0, uint64_t m, uint64_t seed, int dim, const float* cent, int n_cent,
                    float noise_scale, const double* cdf, int n_tickers, float* x, uint32_t* ticker) {
  const uint64_t G1 = 0x9E3779B97F4A7C15ull, G2 = 0xC2B2AE3D27D4EB4Full, G3 = 0x165667B19E3779F9ull;
  const uint64_t s_noise = seed * 0x10001ull + 4ull * 0x9E37ull;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)m; ++i) {
    const uint64_t r = row0 + (uint64_t)i;
    const uint64_t cid = (mix64(r * G2 + (seed + 2)) >> 40) & (uint64_t)(n_cent - 1);
    const double u = (double)(mix64(r * G3 + (seed + 3)) >> 11) * (1.0 / 9007199254740992.0);
    int lo = 0, hi = n_tickers; /* first index with cdf[idx] > u */
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    ticker[i] = (uint32_t)(lo < n_tickers - 1 ? lo : n_tickers - 1);
    const float* c = cent + (size_t)cid * dim;
    float* o = x + (size_t)i * dim;
    for (int j = 0; j < dim; ++j) {
      const float nz = irwin_hall(mix64((r * (uint64_t)dim + (uint64_t)j) * G1 + s_noise)) * noise_scale;
      o[j] = c[j] + nz;
    }
  }
}

/* in-place L2 normalisation of fp32 rows, float32 accumulation in element order (the COSINE collection's
 * normalise-on-insert, oracle/search_oracle.py l2_normalize_f32 up to the summation order) */
void frs_synth_normalize(float* x, uint64_t m, int dim) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)m; ++i) {
    float* o = x + (size_t)i * dim;
    float ss = 0.f;
    for (int j = 0; j < dim; ++j) ss += o[j] * o[j];
    if (ss > 0.f) {
      const float n = __builtin_sqrtf(ss);
      for (int j = 0; j < dim; ++j) o[j] = o[j] / n;
    }
  }
}
