// Oracle helper (TEST INFRASTRUCTURE): the permutation libstdc++'s std::sort produces for
// FeatureDEM's per-region sort -- /root/reference/src/processing/feature_dem.cpp:170 and :230 call
// std::sort(vector<pair<Point2f,float>>, sortbysecdesc) (comparator :6-10, a.second > b.second).
// std::sort is unstable; the tie order matters for parity because the pseudo-Harris scores are coarse.
// Built by oracle/Makefile into oracle/_build/liboracle_helpers.so; this IS libstdc++'s algorithm,
// i.e. what the reference executes (g++/libstdc++ toolchain).
#include <algorithm>
#include <utility>
#include <vector>

extern "C" void oracle_std_sort_desc(const float* score, int n, int* perm) {
  std::vector<std::pair<int, float>> v(n);
  for (int i = 0; i < n; ++i) v[i] = std::make_pair(i, score[i]);
  std::sort(v.begin(), v.end(),
            [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
  for (int i = 0; i < n; ++i) perm[i] = v[i].first;
}
