// Host-side construction of the level schedule (see network.h).
// Behavioural reference: _up_nb / set_dd, wflow/wflow_funcs.py:43-95.
#include "network.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <queue>

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

int64_t label_basins(const Network &net, std::vector<int32_t> &basin, std::vector<int64_t> &cells) {
  const int64_t n = net.ncells, D = net.nlevels();
  basin.assign((size_t)std::max<int64_t>(n, 1), 0);
  cells.clear();
  if (n == 0 || D == 0) return 0;
  const int64_t pb = net.lev_off[(size_t)D - 1], pe = net.lev_off[(size_t)D];
  for (int64_t k = pb; k < pe; k++) basin[(size_t)k] = (int32_t)(k - pb);
  // pits first, then upstream level by level: the children of k are a contiguous run
  for (int64_t k = n - 1; k >= 0; k--) {
    const int32_t b = basin[(size_t)k];
    const int64_t us = net.up_start[(size_t)k];
    for (int j = 0; j < net.nup[(size_t)k]; j++) basin[(size_t)(us + j)] = b;
  }
  cells.assign((size_t)(pe - pb), 0);
  for (int64_t k = 0; k < n; k++) cells[(size_t)basin[(size_t)k]]++;
  return pe - pb;
}

int pack_basins(const std::vector<int64_t> &cells, int ngroups, std::vector<int32_t> &group_of_basin) {
  const int64_t nb = (int64_t)cells.size();
  group_of_basin.assign((size_t)nb, 0);
  if (ngroups < 1) ngroups = 1;
  if (nb <= ngroups) {
    for (int64_t b = 0; b < nb; b++) group_of_basin[(size_t)b] = (int32_t)b;
    return (int)std::max<int64_t>(nb, 1);
  }
  std::vector<int64_t> idx((size_t)nb);
  for (int64_t b = 0; b < nb; b++) idx[(size_t)b] = b;
  std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return cells[(size_t)a] > cells[(size_t)b]; });
  // (load, group) min-heap
  std::priority_queue<std::pair<int64_t, int>, std::vector<std::pair<int64_t, int>>, std::greater<std::pair<int64_t, int>>> heap;
  for (int g = 0; g < ngroups; g++) heap.push({0, g});
  for (int64_t b : idx) {
    auto top = heap.top();
    heap.pop();
    group_of_basin[(size_t)b] = top.second;
    heap.push({top.first + cells[(size_t)b], top.second});
  }
  return ngroups;
}

void regroup(const Network &net, const std::vector<int32_t> &group_of, int G, Network &out, GroupLayout &lay,
             std::vector<int64_t> &newpos) {
  const int64_t n = net.ncells, D = net.nlevels();
  if (G < 1) G = 1;
  lay.ngroups = G;
  // cells per (group, level)
  std::vector<int64_t> cnt((size_t)G * (size_t)std::max<int64_t>(D, 1), 0);
  for (int64_t L = 0; L < D; L++)
    for (int64_t k = net.lev_off[(size_t)L]; k < net.lev_off[(size_t)L + 1]; k++) cnt[(size_t)group_of[(size_t)k] * D + L]++;
  lay.cell_off.assign((size_t)G + 1, 0);
  lay.lev0.assign((size_t)G, (int32_t)D);
  lay.lev_idx.assign((size_t)G + 1, 0);
  lay.lev_off.clear();
  std::vector<int64_t> cursor((size_t)G * (size_t)std::max<int64_t>(D, 1), 0);
  int64_t pos = 0;
  for (int g = 0; g < G; g++) {
    lay.cell_off[(size_t)g] = pos;
    int64_t first = D;
    for (int64_t L = 0; L < D; L++)
      if (cnt[(size_t)g * D + L] > 0) { first = L; break; }
    lay.lev0[(size_t)g] = (int32_t)first;
    lay.lev_idx[(size_t)g] = (int64_t)lay.lev_off.size();
    for (int64_t L = first; L < D; L++) {
      lay.lev_off.push_back((uint32_t)pos);
      cursor[(size_t)g * D + L] = pos;
      pos += cnt[(size_t)g * D + L];
    }
    lay.lev_off.push_back((uint32_t)pos);
  }
  lay.cell_off[(size_t)G] = pos;
  lay.lev_idx[(size_t)G] = (int64_t)lay.lev_off.size();
  newpos.assign((size_t)std::max<int64_t>(n, 1), 0);
  for (int64_t L = 0; L < D; L++)
    for (int64_t k = net.lev_off[(size_t)L]; k < net.lev_off[(size_t)L + 1]; k++)
      newpos[(size_t)k] = cursor[(size_t)group_of[(size_t)k] * D + L]++;
  out.nrow = net.nrow;
  out.ncol = net.ncol;
  out.ncells = n;
  out.lev_off = net.lev_off;  // global level widths (informational: storage is group-major)
  out.order.resize((size_t)n);
  out.up_start.resize((size_t)n);
  out.nup.resize((size_t)n);
  for (int64_t k = 0; k < n; k++) {
    const int64_t q = newpos[(size_t)k];
    out.order[(size_t)q] = net.order[(size_t)k];
    out.nup[(size_t)q] = net.nup[(size_t)k];
    out.up_start[(size_t)q] = net.nup[(size_t)k] ? (uint32_t)newpos[(size_t)net.up_start[(size_t)k]] : 0u;
  }
}

}  // namespace wf
