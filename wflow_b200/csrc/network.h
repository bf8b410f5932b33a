// Drainage-network builder: the level schedule that replaces the reference's serial
// upstream-to-downstream walk (set_dd, wflow/wflow_funcs.py:43-95).
#pragma once
#include <cstdint>
#include <vector>

namespace wf {

// Level-major ("network order") description of one drainage forest.
//
// Cell k (0 <= k < ncells) lives at flat grid index order[k].  Levels are the reference's:
// level 0 holds the cells farthest (in hops) from their pit, the last level the pits; inside
// a level the order is set_dd's discovery order.  Because set_dd appends the upstream
// neighbours of each cell of a level consecutively (in NW,N,NE,W,E,SW,S,SE order), the
// upstream neighbours of cell k are the CONTIGUOUS range [up_start[k], up_start[k]+nup[k]) of the
// previous level, already in the reference's summation order: the inflow gather of the
// kinematic waves is a short contiguous read, no index list needed.
struct Network {
  int nrow = 0, ncol = 0;
  int64_t ncells = 0;
  std::vector<int64_t> lev_off;    // nlevels + 1
  std::vector<int64_t> order;      // ncells, flat grid index
  std::vector<uint32_t> up_start;  // ncells
  std::vector<uint8_t> nup;        // ncells, 0..8
  int64_t nlevels() const { return (int64_t)lev_off.size() - 1; }
};

// Builds the schedule; reproduces set_dd bit-exactly, including its quirk that an upstream
// set consisting solely of flat index 0 is dropped (np.any() on indices, wflow_funcs.py:88).
// ldd: PCRaster codes (1..9 valid, 5 = pit).  Returns 0, or -1 if the ldd has a cycle.
int build_network(const uint8_t *ldd, int nrow, int ncol, Network &net);

}  // namespace wf
