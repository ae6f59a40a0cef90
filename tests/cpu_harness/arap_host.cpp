// Host build of dimo_b200/csrc/arap_math.h for the CPU tests (tests/test_arap_math_cpu.py): the per-vertex ARAP
// arithmetic the CUDA kernels run, compiled with g++ and driven through ctypes.  Test infrastructure, not product.
#include "../../dimo_b200/csrc/arap_math.h"

extern "C" {

// nodes [T,M,3], nbr [M,K], mult [M] or NULL -> *energy, grad [T,M,3] (zeroed here)
void arap_energy_host(int T, int M, int K, const float* nodes, const int64_t* nbr, const float* mult, double* energy,
                      float* grad) {
  for (int64_t e = 0; e < (int64_t)T * M * 3; ++e) grad[e] = 0.0f;
  auto add = [](float* p, float v) { *p += v; };
  double total = 0.0;
  for (int t = 1; t < T; ++t)
    for (int i = 0; i < M; ++i) {
      const float m = mult ? mult[i] : 1.0f;
      if (m == 0.0f) continue;
      total += dimo::arap::vertex_term(K, nodes, nodes + (int64_t)t * M * 3, nbr + (int64_t)i * K, i, m, grad,
                                       grad + (int64_t)t * M * 3, add);
    }
  *energy = total;
}

void arap_connectivity_host(int T, int M, int Kq, int K, float radius, const float* nodes, int64_t* nbr, int* count) {
  for (int i = 0; i < M; ++i)
    count[i] = dimo::arap::common_neighbours(T, M, Kq, K, radius * radius, nodes, i, nbr + (int64_t)i * K);
}

void arap_rotation_host(const double* S9, double* R9) {
  double S[3][3], R[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) S[a][b] = S9[3 * a + b];
  dimo::arap::rotation_from_covariance(S, R);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) R9[3 * a + b] = R[a][b];
}
}
