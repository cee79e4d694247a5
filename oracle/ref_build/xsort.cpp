// Test infrastructure (oracle/): the permutation libstdc++'s std::sort applies for the comparator of
// TrackManager::manageTracks (track_manager.cpp:274-276: descending track length, NOT stable).  oracle/track_manager.py
// delegates the sort to the standard library so that the order among equal lengths is the reference's.
#include <algorithm>
#include <vector>
struct Item { int len, idx; };
extern "C" void xsort_desc_by_len(const int* len, int n, int* perm) {
  std::vector<Item> v(n);
  for (int i = 0; i < n; ++i) v[i] = {len[i], i};
  std::sort(v.begin(), v.end(), [](const Item& a, const Item& b) { return a.len > b.len; });
  for (int i = 0; i < n; ++i) perm[i] = v[i].idx;
}
