// Host-side replay of the reference's greedy region selection (score/sv_level/LiDAL.py:242-270 and :293-325).
//
// The walk is inherently sequential (every decision depends on the regions accepted so far), so it is not a kernel: it is
// the native host half of the selection step, fed by the device-built 5 m neighbour lists (lb_region_pairs) and the device
// argsort.  The one subtle part is the reference's tie rule: it iterates a CPython `set` of the accepted regions and takes
// the FIRST member within range, i.e. the member with the lowest slot in the set's hash table.  SetModel replays that
// table (Objects/setobject.c: open addressing, 9 linear probes, perturb shift 5, growth at fill*5 >= mask*3 to the first
// power of two above used*4, dummies left by removals are reused by later insertions) for small non-negative int keys,
// for which hash(key) == key.  The Python caller verifies the model against the running interpreter's own set before it
// uses this entry point (lidal_b200.score._set_model) and replays the walk on a real set otherwise.
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace lb {
namespace {

struct SetModel {
  static constexpr int64_t UNUSED = -2, DUMMY = -1;
  std::vector<int64_t> keys;
  std::vector<int64_t> slot;     // region id -> slot in `keys`, -1 when not a member
  int64_t mask = 7, fill = 0, used = 0;
  bool last_dummy;

  SetModel(int64_t n_ids, bool last_dummy_) : keys(8, UNUSED), slot((size_t)n_ids, -1), last_dummy(last_dummy_) {}

  static int64_t clean_slot(const std::vector<int64_t>& tab, int64_t mask, int64_t key) {
    uint64_t perturb = (uint64_t)key;
    int64_t i = key & mask;
    for (;;) {
      if (tab[i] == UNUSED) return i;
      if (i + 9 <= mask)
        for (int64_t j = i + 1; j <= i + 9; ++j)
          if (tab[j] == UNUSED) return j;
      perturb >>= 5;
      i = (int64_t)(((uint64_t)i * 5 + 1 + perturb) & (uint64_t)mask);
    }
  }

  void resize(int64_t minused) {
    int64_t size = 8;
    while (size <= minused) size <<= 1;
    std::vector<int64_t> fresh((size_t)size, UNUSED);
    const int64_t m = size - 1;
    for (int64_t key : keys) {                      // old table order, clean insertion
      if (key < 0) continue;
      const int64_t s = clean_slot(fresh, m, key);
      fresh[s] = key;
      slot[key] = s;
    }
    keys.swap(fresh);
    mask = m;
    fill = used;
  }

  void add(int64_t key) {
    uint64_t perturb = (uint64_t)key;
    int64_t i = key & mask, free_slot = -1;
    for (;;) {
      const int64_t last = (i + 9 <= mask) ? i + 9 : i;
      for (int64_t j = i; j <= last; ++j) {
        const int64_t k = keys[j];
        if (k == UNUSED) {
          int64_t at = j;
          if (free_slot >= 0) at = free_slot; else ++fill;
          keys[at] = key;
          slot[key] = at;
          ++used;
          if (free_slot < 0 && fill * 5 >= mask * 3) resize(used > 50000 ? used * 2 : used * 4);
          return;
        }
        if (k == key) return;
        if (k == DUMMY && (last_dummy || free_slot < 0)) free_slot = j;
      }
      perturb >>= 5;
      i = (int64_t)(((uint64_t)i * 5 + 1 + perturb) & (uint64_t)mask);
    }
  }

  void remove(int64_t key) {
    keys[slot[key]] = DUMMY;
    slot[key] = -1;
    --used;
  }
};

}  // namespace
}  // namespace lb

extern "C" int lb_select_walk(const int64_t* visit, int64_t n_visit, const double* interds, const double* interes,
                              const int64_t* pnums, const int64_t* row_ptr, const int32_t* nbr_idx, int64_t n_regions,
                              int64_t* flags, int64_t flag_value, int64_t point_limit, int prefer_higher_entropy, int skip_zero,
                              int set_last_dummy, int64_t* n_added_out) {
  using namespace lb;
  if (!visit || !interds || !interes || !pnums || !row_ptr || !flags || n_regions < 0 || (n_regions > 0 && !nbr_idx && row_ptr[n_regions] > 0)) {
    set_error("lb_select_walk: null argument");
    return LB_EINVAL;
  }
  SetModel added(n_regions, set_last_dummy != 0);
  for (int64_t v = 0; v < n_visit; ++v) {
    const int64_t sv = visit[v];
    if (sv < 0 || sv >= n_regions) { set_error("lb_select_walk: region id out of range"); return LB_EINVAL; }
    if (skip_zero && interds[sv] == 0.0) continue;
    int64_t hit = -1, hit_slot = INT64_MAX;
    if (added.used > 0)
      for (int64_t q = row_ptr[sv]; q < row_ptr[sv + 1]; ++q) {
        const int64_t s = added.slot[nbr_idx[q]];
        if (s >= 0 && s < hit_slot) { hit_slot = s; hit = nbr_idx[q]; }     // first in-range member in the set's iteration order
      }
    if (hit >= 0) {
      const bool better = prefer_higher_entropy ? interes[hit] < interes[sv] : interes[hit] > interes[sv];
      if (better) {
        flags[sv] = flag_value;
        flags[hit] = 0;
        added.add(sv);
        added.remove(hit);
        point_limit += pnums[hit] - pnums[sv];
      }
      continue;
    }
    point_limit -= pnums[sv];
    if (point_limit < 0) break;
    flags[sv] = flag_value;
    added.add(sv);
  }
  if (n_added_out) *n_added_out = added.used;
  return LB_OK;
}
