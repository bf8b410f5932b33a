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

// ---- storage layout: independent groups of whole basins -----------------------------------------
// Every dependency of the lateral flow follows the LDD downstream, so the trees of the drainage
// forest (one per pit, "basins") never exchange anything.  The device arrays are therefore stored
// group-major: a group is a set of whole basins that one thread-block cluster sweeps on its own
// (its barriers are cluster barriers, not grid barriers); inside a group the cells are level-major
// with the reference's in-level order, so the upstream neighbours of a cell remain one contiguous
// run of the group's previous level, in the reference's summation order.
struct GroupLayout {
  int ngroups = 0;
  std::vector<int64_t> cell_off;   // ngroups + 1: storage range of each group
  std::vector<int32_t> lev0;       // ngroups: first (global) level that has cells of the group
  std::vector<int64_t> lev_idx;    // ngroups + 1: range of the group in lev_off
  std::vector<uint32_t> lev_off;   // per group: (nlevels - lev0 + 1) absolute storage offsets
};

// Basin of every cell (network order): the rank of its pit within the last level.
// Returns the number of basins; cells[b] = size of basin b.
int64_t label_basins(const Network &net, std::vector<int32_t> &basin, std::vector<int64_t> &cells);

// Longest-processing-time packing of basins into at most `ngroups` groups of similar cell count.
// Returns the number of groups actually used; group_of_basin[b] in [0, that).
int pack_basins(const std::vector<int64_t> &cells, int ngroups, std::vector<int32_t> &group_of_basin);

// Re-order `net` group-major.  group_of[k] = group of the cell at network position k.
// out: the same forest in storage order; newpos[k] = storage position of network position k.
void regroup(const Network &net, const std::vector<int32_t> &group_of, int ngroups, Network &out, GroupLayout &lay,
             std::vector<int64_t> &newpos);

}  // namespace wf
