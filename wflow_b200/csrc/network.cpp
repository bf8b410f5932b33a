// Host-side construction of the level schedule (see network.h).
// Behavioural reference: _up_nb / set_dd, wflow/wflow_funcs.py:43-95.
#include "network.h"

#include <cstring>

namespace wf {

int build_network(const uint8_t *ldd, int nrow, int ncol, Network &net) {
  const int64_t N = (int64_t)nrow * ncol;
  int64_t nvalid = 0;
  for (int64_t i = 0; i < N; i++) nvalid += (ldd[i] >= 1 && ldd[i] <= 9);

  // Downstream-first breadth-first storage: [pits][1 hop][2 hops]...  The children of the
  // cell at position p are appended consecutively, so they form one contiguous run.
  std::vector<int64_t> cells((size_t)nvalid);
  std::vector<int64_t> cstart((size_t)nvalid);
  std::vector<uint8_t> ccnt((size_t)nvalid);
  std::vector<int64_t> loff;
  int64_t n = 0;
  loff.push_back(0);
  for (int64_t i = 0; i < N; i++)
    if (ldd[i] == 5) cells[n++] = i;  // np.where(ldd == pit) in flat order
  // code an upstream neighbour at window offset (dr,dc) must carry to drain into the centre
  static const uint8_t us[3][3] = {{3, 2, 1}, {6, 5, 4}, {9, 8, 7}};
  int64_t beg = 0;
  do {  // the reference always emits the pit level, even when it is empty
    const int64_t end = n;
    loff.push_back(end);
    for (int64_t p = beg; p < end; p++) {
      const int64_t idx0 = cells[p];
      const int64_t r = idx0 / ncol, c = idx0 - r * ncol;
      int64_t up[8];
      int cnt = 0;
      bool any_nonzero = false;
      for (int dr = -1; dr <= 1; dr++) {
        const int64_t row = r + dr;
        if (row < 0 || row >= nrow) continue;
        for (int dc = -1; dc <= 1; dc++) {
          if (dr == 0 && dc == 0) continue;
          const int64_t col = c + dc;
          if (col < 0 || col >= ncol) continue;
          const int64_t idx = idx0 + dc + (int64_t)dr * ncol;
          if (ldd[idx] == us[dr + 1][dc + 1]) {
            up[cnt++] = idx;
            any_nonzero |= (idx != 0);
          }
        }
      }
      cstart[p] = n;
      if (!any_nonzero) cnt = 0;  // reference quirk: `if np.any(idx_up)` tests VALUES
      if (n + cnt > nvalid) return -1;  // cannot happen on a forest
      for (int j = 0; j < cnt; j++) cells[n++] = up[j];
      ccnt[p] = (uint8_t)cnt;
    }
    beg = end;
  } while (beg < n);
  const int64_t D = (int64_t)loff.size() - 1;
  net.nrow = nrow;
  net.ncol = ncol;
  net.ncells = n;
  net.lev_off.assign((size_t)D + 1, 0);
  net.order.resize((size_t)n);
  net.up_start.resize((size_t)n);
  net.nup.resize((size_t)n);
  // reverse the level order: new level L = D-1-d
  std::vector<int64_t> newoff((size_t)D + 1, 0);
  for (int64_t L = 0; L < D; L++) {
    const int64_t d = D - 1 - L;
    newoff[L + 1] = newoff[L] + (loff[d + 1] - loff[d]);
  }
  net.lev_off = newoff;
  for (int64_t d = 0; d < D; d++) {
    const int64_t L = D - 1 - d;
    const int64_t w = loff[d + 1] - loff[d];
    std::memcpy(&net.order[(size_t)newoff[L]], &cells[(size_t)loff[d]], sizeof(int64_t) * (size_t)w);
    for (int64_t q = 0; q < w; q++) {
      const int64_t p = loff[d] + q;
      const int64_t k = newoff[L] + q;
      net.nup[(size_t)k] = ccnt[(size_t)p];
      // children live in old level d+1 = new level L-1, same relative position
      int64_t us_new = 0;
      if (ccnt[(size_t)p] > 0) us_new = newoff[L - 1] + (cstart[(size_t)p] - loff[d + 1]);
      net.up_start[(size_t)k] = (uint32_t)us_new;
    }
  }
  return 0;
}

}  // namespace wf
