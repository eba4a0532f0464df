// Golden vectors for frames.std_shuffle (tests/test_host_logic.py): std::shuffle of iota(n) with std::mt19937 seeded like
// ReprojectionFactor's keypoint draw (core/gtsam/reprojection_factor.cpp:43-50).  Build: g++ -O2 make_std_shuffle.cpp
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>
int main()
{
  const long cases[][2] = {{10, 0}, {11, 6}, {1000, 42}, {65535, 3}, {65536, 3}, {81920, 12}, {12288, 2}};
  printf("# gcc %d.%d.%d\n", __GNUC__, __GNUC_MINOR__, __GNUC_PATCHLEVEL__);
  for (auto &c : cases)
  {
    std::vector<long> idx(c[0]);
    std::iota(idx.begin(), idx.end(), 0);
    std::mt19937 g;
    g.seed(c[1]);
    std::shuffle(idx.begin(), idx.end(), g);
    printf("%ld %ld", c[0], c[1]);
    for (int k = 0; k < 16 && k < c[0]; ++k)
      printf(" %ld", idx[k]);
    unsigned long h = 1469598103934665603ul;
    for (long v : idx)
      h = (h ^ (unsigned long)v) * 1099511628211ul;
    printf(" | %lu\n", h);
  }
}
