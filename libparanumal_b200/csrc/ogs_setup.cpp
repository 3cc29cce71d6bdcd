// Host-side gather-scatter setup: shared-node discovery, unique-owner flagging, local/halo
// CSR maps and the pairwise exchange lists.
//
// Behaviour follows ogsBase_t::Setup (libs/ogs/ogsSetup.cpp:69-190) and the ogsPairwise_t
// constructor (libs/ogs/ogsPairwise.cpp:194-415) so that every map is bit-identical to the
// reference's for the same ids in the same call sequence:
//   * the owner copy of each id group is the (rand() % groupsize)-th record of the group in the
//     order left by an UNSTABLE std::sort keyed on |id| (ogsSetup.cpp:245-275) -> we call the same
//     libstdc++ std::sort with an equivalent comparator and the same glibc rand() stream;
//   * gathered rows are numbered by first appearance in local order, owner rows first
//     (ogsSetup.cpp:411-433); columns ascend in local id (ogsSetup.cpp:637-663).
// Collectives go through libp_comm_s (host callbacks; memcpy when size==1).
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <numeric>

#include "ogs.hpp"

using namespace libp_b200;

namespace {

struct Node {
  dlong localId;  // local node id
  hlong baseId;   // signed global id
  dlong newId;    // scratch / gathered row id
  int sign;       // +-1 local group, +-2 group shared between ranks (negative: no owner copy here)
  int rank;       // origin rank
  int destRank;   // rendezvous rank = |id| % size
};

inline hlong habs(hlong v) { return v < 0 ? -v : v; }

// std::vector without value-initialisation: `NodeVec out(n)` neither zero-fills 4 GB on one thread nor touches the
// pages - the first touch happens inside the parallel loops that fill it
template <class T>
struct default_init_alloc : std::allocator<T> {
  template <class U> struct rebind { typedef default_init_alloc<U> other; };
  default_init_alloc() = default;
  template <class U> default_init_alloc(const default_init_alloc<U>&) {}
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(args) == 0) ::new (static_cast<void*>(p)) U;
    else ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
};
typedef std::vector<Node, default_init_alloc<Node>> NodeVec;

// LIBP_OGS_TIMING=1: wall-clock laps of the setup stages on stderr
struct SubLap {
  bool on;
  std::chrono::steady_clock::time_point t;
  SubLap() : on(getenv("LIBP_OGS_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
  void operator()(const char* what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[ogs setup]     %-32s %.3f s\n", what, std::chrono::duration<double>(n - t).count());
    t = n;
  }
};

// ---- std::sort, run in parallel, with the identical result ------------------------------------------------------
// The owner copy of an id group depends on the order an UNSTABLE std::sort leaves the group in (ogsSetup.cpp:245-275),
// so the sort cannot be swapped for another algorithm.  libstdc++'s std::sort is introsort: a loop that partitions
// around a median-of-three pivot, recurses into the right part and continues with the left one, heap-sorts a range
// when the depth limit 2*floor(log2 n) is hit, and finishes with one insertion-sort pass over ranges of <= 16
// elements.  The two parts of a partition never interact, so running the recursion as OpenMP tasks executes exactly
// the same comparisons and swaps on every range - the output equals std::sort's bit for bit (libp_ogs_sort_selftest
// checks that against std::sort itself).  The final pass is done per <= 16-element leaf: an element never moves out
// of its leaf in the sequential pass either (everything left of a leaf compares not-greater).
// Only comparison outcomes steer the algorithm, so sorting compact (key, index) pairs gives the permutation that
// sorting the 32-byte records by the same key would give.
struct KeyIdx { unsigned long long key; dlong idx; };

struct ExactSort {
  KeyIdx* base;
  static bool lt(const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; }
  static void move_median_to_first(KeyIdx* result, KeyIdx* a, KeyIdx* b, KeyIdx* c) {
    if (lt(*a, *b)) {
      if (lt(*b, *c)) std::iter_swap(result, b);
      else if (lt(*a, *c)) std::iter_swap(result, c);
      else std::iter_swap(result, a);
    } else if (lt(*a, *c)) std::iter_swap(result, a);
    else if (lt(*b, *c)) std::iter_swap(result, c);
    else std::iter_swap(result, b);
  }
  static KeyIdx* unguarded_partition(KeyIdx* first, KeyIdx* last, KeyIdx* pivot) {
    while (true) {
      while (lt(*first, *pivot)) ++first;
      --last;
      while (lt(*pivot, *last)) --last;
      if (!(first < last)) return first;
      std::iter_swap(first, last);
      ++first;
    }
  }
  static void insertion_sort(KeyIdx* first, KeyIdx* last) {  // what the final pass does inside one leaf
    for (KeyIdx* i = first + 1; i < last; ++i) {
      KeyIdx val = *i;
      KeyIdx* pos = i;
      while (pos > first && lt(val, *(pos - 1))) { *pos = *(pos - 1); --pos; }
      *pos = val;
    }
  }
  static void loop(KeyIdx* first, KeyIdx* last, long depth) {
    while (last - first > 16) {
      if (depth == 0) {
        std::partial_sort(first, last, last, lt);  // the heap-sort fallback of std::__introsort_loop
        return;
      }
      --depth;
      KeyIdx* mid = first + (last - first) / 2;
      move_median_to_first(first, first + 1, mid, last - 1);
      KeyIdx* cut = unguarded_partition(first + 1, last, first);
      if (last - cut > (1 << 15)) {
#pragma omp task default(none) firstprivate(cut, last, depth)
        loop(cut, last, depth);
      } else {
        loop(cut, last, depth);
      }
      last = cut;
    }
    if (last - first > 1) insertion_sort(first, last);
  }
  static void sort(KeyIdx* first, size_t n) {
    if (n < 2) return;
    long lg = 0;
    for (size_t m = n; m > 1; m >>= 1) ++lg;
#pragma omp parallel
#pragma omp single nowait
    loop(first, first + n, 2 * lg);
  }
};

// ---- glibc rand() in bulk ----------------------------------------------------------------------------------------
// ogsBase_t::FindSharedNodes draws one rand() per id group (ogsSetup.cpp:262) from the stream the embedding program
// seeded; 89 M calls through rand()'s lock cost close to a second at 64^3.  glibc's default generator (TYPE_3 of
// random_r: r[i] = r[i-3] + r[i-31], output r[i] >> 1) keeps its table in a state array that initstate()/setstate()
// hand out, with the rear index saved in word 0 (5 * rear + type) whenever the state is switched.  take() switches to
// a scratch state, advances the REAL table n steps inline, and switches back: the values are the ones n rand() calls
// would have returned and the process-wide stream continues exactly where they would have left it.  Any other
// generator type (a program that called initstate itself) falls back to calling rand().
struct GlibcRandBulk {
  static void take(std::vector<int, default_init_alloc<int>>& out, size_t n) {
    out.resize(n);
    if (n == 0) return;
    alignas(8) static int32_t scratch[32];
    char* oldp = initstate(1u, reinterpret_cast<char*>(scratch), 128);
    int32_t* st = reinterpret_cast<int32_t*>(oldp);
    if (oldp == nullptr || st[0] % 5 != 3 || st[0] < 0 || st[0] / 5 >= 31) {
      if (oldp != nullptr) setstate(oldp);
      for (size_t i = 0; i < n; ++i) out[i] = rand();
      return;
    }
    int rear = st[0] / 5, front = (rear + 3) % 31;
    uint32_t tbl[31];
    for (int i = 0; i < 31; ++i) tbl[i] = (uint32_t)st[1 + i];
    for (size_t i = 0; i < n; ++i) {
      const uint32_t v = (tbl[front] += tbl[rear]);
      out[i] = (int)(v >> 1);
      if (++front == 31) front = 0;
      if (++rear == 31) rear = 0;
    }
    for (int i = 0; i < 31; ++i) st[1 + i] = (int32_t)tbl[i];
    st[0] = 5 * rear + 3;
    setstate(oldp);
  }
};

// ---- records sorted by |baseId|: cut into chunks that start at group boundaries --------------------------------
struct GroupChunks {
  std::vector<size_t> begin;   // [nchunks + 1] record index, every entry is the first record of a group (or n)
  std::vector<size_t> group0;  // [nchunks + 1] index of the chunk's first group
  size_t ngroups = 0;
  GroupChunks(const NodeVec& a, int nchunks) : begin((size_t)nchunks + 1), group0((size_t)nchunks + 1, 0) {
    const size_t n = a.size(), per = (n + nchunks - 1) / std::max(nchunks, 1);
    for (int c = 0; c <= nchunks; ++c) {
      size_t b = std::min(n, (size_t)c * per);
      while (b > 0 && b < n && habs(a[b].baseId) == habs(a[b - 1].baseId)) ++b;
      begin[(size_t)c] = b;
    }
    begin[(size_t)nchunks] = n;
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nchunks; ++c) {
      size_t g = 0;
      for (size_t i = begin[(size_t)c]; i < begin[(size_t)c + 1]; ++i)
        g += (i == begin[(size_t)c]) || habs(a[i].baseId) != habs(a[i - 1].baseId);
      group0[(size_t)c + 1] = g;
    }
    for (int c = 0; c < nchunks; ++c) group0[(size_t)c + 1] += group0[(size_t)c];
    ngroups = group0[(size_t)nchunks];
  }
};
constexpr int kSetupChunks = 1024;

// a = a permuted so that it is sorted by key(a[i]) with std::sort's exact tie order
template <class KeyFn>
void exact_sort_nodes(std::vector<struct Node>& a, KeyFn key);

template <class KeyFn>
void exact_sort_nodes(NodeVec& a, KeyFn key) {
  const size_t n = a.size();
  std::vector<KeyIdx, default_init_alloc<KeyIdx>> k(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) { k[i].key = key(a[i]); k[i].idx = (dlong)i; }
  const bool timing = getenv("LIBP_OGS_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  ExactSort::sort(k.data(), n);
  if (timing)
    fprintf(stderr, "[ogs setup]   exact sort of %zu keys: %.3f s\n", n,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  NodeVec out(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) out[i] = a[(size_t)k[i].idx];
  a.swap(out);
}

// out[key(in[n])] = in[n]
template <class Key>
void permute_by(NodeVec& a, Key key) {
  NodeVec out(a.size());
  const size_t n = a.size();
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) out[(size_t)key(a[i])] = a[i];
  a.swap(out);
}

void exchange_nodes(const libp_comm_s& comm, const NodeVec& send, const std::vector<int>& sendCounts,
                    NodeVec& recv, std::vector<int>& recvCounts) {
  const int size = comm.size;
  if (size == 1) {  // one rank: the exchange is a copy
    recvCounts = sendCounts;
    recv.resize(send.size());
    const size_t n = send.size();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) recv[i] = send[i];
    return;
  }
  recvCounts.assign(size, 0);
  comm.alltoall(sendCounts.data(), recvCounts.data(), sizeof(int));
  std::vector<int64_t> sc(size), so(size), rc(size), ro(size);
  int64_t s = 0, r = 0;
  for (int i = 0; i < size; ++i) {
    sc[i] = (int64_t)sendCounts[i] * (int64_t)sizeof(Node);
    so[i] = s;
    s += sc[i];
    rc[i] = (int64_t)recvCounts[i] * (int64_t)sizeof(Node);
    ro[i] = r;
    r += rc[i];
  }
  recv.resize((size_t)(r / (int64_t)sizeof(Node)));
  comm.alltoallv(send.data(), sc.data(), so.data(), recv.data(), rc.data(), ro.data());
}

// Flag, for every id group, the owner copy (when unique) and whether the group spans ranks.
void find_shared_nodes(libp_ogs_s& o, NodeVec& nodes) {
  const libp_comm_s& comm = *o.comm;
  const int size = comm.size;
  std::vector<int> sendCounts(size, 0), sendOffsets(size + 1, 0);
  NodeVec recv;
  std::vector<int> recvCounts;
  if (size == 1) {
    // one rank: send order = local order and the exchange is the identity - no counting, no permutation, no copy
    sendCounts[0] = (int)nodes.size();
    recvCounts = sendCounts;
    recv.swap(nodes);
  } else {
    for (const Node& n : nodes) sendCounts[n.destRank]++;
    for (int r = 0; r < size; ++r) sendOffsets[r + 1] = sendOffsets[r] + sendCounts[r];
    {
      std::vector<int> fill(size, 0);
      for (Node& n : nodes) n.newId = sendOffsets[n.destRank] + fill[n.destRank]++;
    }
    permute_by(nodes, [](const Node& n) { return n.newId; });  // send order: by destRank, local order inside
    exchange_nodes(comm, nodes, sendCounts, recv, recvCounts);
  }
  SubLap sub;
  const dlong recvN = (dlong)recv.size();
#pragma omp parallel for schedule(static)
  for (dlong n = 0; n < recvN; ++n) recv[n].newId = n;

  // same algorithm + equivalent strict weak order as the reference => same tie order
  exact_sort_nodes(recv, [](const Node& a) { return (unsigned long long)habs(a.baseId); });
  sub("find: sort by |id|");

  // one pass per id group, groups dealt to threads in chunks; the rand() draws are taken from the stream up front,
  // one per group in sorted order, exactly as the sequential loop of the reference consumes them
  int is_unique = 1;
  {
    const GroupChunks gc(recv, kSetupChunks);
    sub("find: group chunks");
    std::vector<int, default_init_alloc<int>> draws;
    if (o.unique) GlibcRandBulk::take(draws, gc.ngroups);
    sub("find: rand() draws");
    long long badId = 0;
    int badCount = -1;
#pragma omp parallel for schedule(dynamic, 4) reduction(min : is_unique)
    for (int c = 0; c < kSetupChunks; ++c) {
      size_t g = gc.group0[(size_t)c];
      dlong start = (dlong)gc.begin[(size_t)c];
      const dlong cend = (dlong)gc.begin[(size_t)c + 1];
      for (dlong n = start; n < cend; ++n) {
        if (n == cend - 1 || habs(recv[n].baseId) != habs(recv[n + 1].baseId)) {
          const dlong end = n + 1;
          int positiveCount = 0;
          if (o.unique) {
            const hlong baseId = habs(recv[start].baseId);
            const int m = draws[g] % (end - start);
            for (dlong i = start; i < end; ++i) recv[i].baseId = -baseId;
            recv[start + m].baseId = baseId;
            positiveCount = 1;
          } else {
            for (dlong i = start; i < end; ++i)
              if (recv[i].baseId > 0) positiveCount++;
            if (positiveCount != 1) is_unique = 0;
          }
          if (o.kind == LIBP_HALO && positiveCount != 1) {
#pragma omp critical(libp_ogs_bad_halo)
            { badId = (long long)habs(recv[start].baseId); badCount = positiveCount; }
          }
          int shared = 1;
          const int r0 = recv[start].rank;
          for (dlong i = start + 1; i < end; ++i)
            if (recv[i].rank != r0) { shared = 2; break; }
          for (dlong i = start; i < end; ++i) recv[i].sign = shared;
          start = end;
          ++g;
        }
      }
    }
    LIBP_CHECK(badCount < 0, "Found " + std::to_string(badCount) + " positive Ids for baseId: " + std::to_string(badId) + ".");
  }
  sub("find: owner / shared flags");
  int64_t u = is_unique;
  comm.allreduce_i64(&u, 1, LIBP_MIN);
  o.gather_defined = (u == 1);

  permute_by(recv, [](const Node& n) { return n.newId; });  // back to arrival order
  sub("find: back to arrival order");
  if (size == 1) {
    nodes.swap(recv);
    return;
  }
  NodeVec back;
  std::vector<int> backCounts;
  exchange_nodes(comm, recv, recvCounts, back, backCounts);
  nodes.swap(back);  // now in send order again, with signs / owner flags filled in
}

// Number the gathered rows and collect, per shared id, who else holds it.
void construct_shared_nodes(libp_ogs_s& o, NodeVec& nodes, NodeVec& sharedNodes) {
  const libp_comm_s& comm = *o.comm;
  const int size = comm.size;
  const dlong Nids = (dlong)nodes.size();

  SubLap sub;
  // |id| ascending, the owner (positive) copy first inside a group: one integer key
  exact_sort_nodes(nodes, [](const Node& a) {
    return ((unsigned long long)habs(a.baseId) << 1) | (unsigned long long)(a.baseId < 0 ? 1 : 0);
  });

  sub("construct: sort by (|id|, owner)");
  // group pass (parallel over chunks of whole groups): ownership sign, group number, first local appearance
  const GroupChunks gc(nodes, kSetupChunks);
  const dlong NbaseIds = (dlong)gc.ngroups;
  std::vector<dlong, default_init_alloc<dlong>> firstPos((size_t)NbaseIds);
  std::vector<signed char, default_init_alloc<signed char>> groupSign((size_t)NbaseIds);
  std::vector<NodeVec> chunkShared((size_t)kSetupChunks);
  long long nLT = 0, nLP = 0, nHT = 0, nHP = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nLT, nLP, nHT, nHP)
  for (int c = 0; c < kSetupChunks; ++c) {
    dlong g = (dlong)gc.group0[(size_t)c];
    dlong start = (dlong)gc.begin[(size_t)c];
    const dlong cend = (dlong)gc.begin[(size_t)c + 1];
    for (dlong n = start; n < cend; ++n) {
      if (n == cend - 1 || habs(nodes[n].baseId) != habs(nodes[n + 1].baseId)) {
        const dlong end = n + 1;
        int sign = std::abs(nodes[start].sign);
        if (nodes[start].baseId < 0) {
          sign = -sign;
          for (dlong i = start; i < end; ++i) nodes[i].sign = sign;
        }
        if (std::abs(sign) == 1) { nLT++; if (sign == 1) nLP++; }
        else { nHT++; if (sign == 2) nHP++; }
        dlong fp = nodes[start].localId;
        for (dlong i = start; i < end; ++i) { nodes[i].newId = g; fp = std::min(fp, nodes[i].localId); }
        firstPos[(size_t)g] = fp;
        groupSign[(size_t)g] = (signed char)sign;
        if (std::abs(sign) == 2) chunkShared[(size_t)c].push_back(nodes[start]);
        ++g;
        start = end;
      }
    }
  }
  o.NlocalT = (dlong)nLT; o.NlocalP = (dlong)nLP; o.NhaloT = (dlong)nHT; o.NhaloP = (dlong)nHP;
  o.Ngather = o.NlocalP + o.NhaloP;
  int64_t ng = o.Ngather;
  comm.allreduce_i64(&ng, 1, LIBP_ADD);
  o.NgatherGlobal = ng;

  NodeVec sendShared;
  sendShared.reserve((size_t)o.NhaloT);
  for (const NodeVec& v : chunkShared) sendShared.insert(sendShared.end(), v.begin(), v.end());
  chunkShared.clear();

  sub("construct: group pass");
  permute_by(nodes, [](const Node& n) { return n.localId; });  // compressed local order
  sub("construct: back to local order");

  // renumber groups by first appearance: owner-local, other-local, owner-halo, other-halo.  A group's new number is
  // the count of groups of its class that appear earlier in local order = a prefix sum over the first appearances.
  std::vector<dlong, default_init_alloc<dlong>> indexMap((size_t)NbaseIds);
  {
    auto cls = [](int sign) { return sign == 1 ? 0 : sign == -1 ? 1 : sign == 2 ? 2 : 3; };
    const int nch = kSetupChunks;
    const size_t per = ((size_t)Nids + nch - 1) / nch;
    std::vector<std::array<dlong, 4>> cnt((size_t)nch + 1, std::array<dlong, 4>{0, 0, 0, 0});
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nch; ++c) {
      std::array<dlong, 4> k{0, 0, 0, 0};
      const size_t b = std::min<size_t>((size_t)Nids, c * per), e = std::min<size_t>((size_t)Nids, b + per);
      for (size_t n = b; n < e; ++n) {
        const dlong g = nodes[n].newId;
        if (firstPos[(size_t)g] == (dlong)n) k[(size_t)cls(groupSign[(size_t)g])]++;
      }
      cnt[(size_t)c + 1] = k;
    }
    cnt[0] = {0, o.NlocalP, 0, o.NhaloP};
    for (int c = 0; c < nch; ++c)
      for (int q = 0; q < 4; ++q) cnt[(size_t)c + 1][(size_t)q] += cnt[(size_t)c][(size_t)q];
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nch; ++c) {
      std::array<dlong, 4> k = cnt[(size_t)c];
      const size_t b = std::min<size_t>((size_t)Nids, c * per), e = std::min<size_t>((size_t)Nids, b + per);
      for (size_t n = b; n < e; ++n) {
        const dlong g = nodes[n].newId;
        if (firstPos[(size_t)g] == (dlong)n) indexMap[(size_t)g] = k[(size_t)cls(groupSign[(size_t)g])]++;
      }
    }
#pragma omp parallel for schedule(static)
    for (dlong n = 0; n < Nids; ++n) nodes[n].newId = indexMap[(size_t)nodes[n].newId];
  }
  sub("construct: first-appearance renumbering");
  for (Node& s : sendShared) s.localId = indexMap[s.newId];

  std::sort(sendShared.begin(), sendShared.end(), [](const Node& a, const Node& b) { return a.destRank < b.destRank; });
  std::vector<int> sendCounts(size, 0);
  for (const Node& s : sendShared) sendCounts[s.destRank]++;
  NodeVec recvShared;
  std::vector<int> recvCounts;
  exchange_nodes(comm, sendShared, sendCounts, recvShared, recvCounts);

  std::sort(recvShared.begin(), recvShared.end(),
            [](const Node& a, const Node& b) { return habs(a.baseId) < habs(b.baseId); });

  const dlong recvN = (dlong)recvShared.size();
  std::vector<int> shCounts(size, 0), shOffsets(size + 1, 0);
  dlong start = 0;
  for (dlong n = 0; n < recvN; ++n)
    if (n == recvN - 1 || habs(recvShared[n].baseId) != habs(recvShared[n + 1].baseId)) {
      const dlong end = n + 1;
      for (dlong i = start; i < end; ++i) shCounts[recvShared[i].rank] += end - start - 1;
      start = end;
    }
  for (int r = 0; r < size; ++r) shOffsets[r + 1] = shOffsets[r] + shCounts[r];
  NodeVec shSend((size_t)shOffsets[size]);
  std::vector<int> fill(size, 0);
  start = 0;
  for (dlong n = 0; n < recvN; ++n)
    if (n == recvN - 1 || habs(recvShared[n].baseId) != habs(recvShared[n + 1].baseId)) {
      const dlong end = n + 1;
      for (dlong i = start; i < end; ++i) {
        const int r = recvShared[i].rank;
        dlong sid = shOffsets[r] + fill[r];
        for (dlong j = start; j < end; ++j) {
          if (j == i) continue;
          shSend[sid] = recvShared[j];       // the other participant (its rank, its gathered row in localId)
          shSend[sid].newId = recvShared[i].localId;  // receiver's own gathered row for this id
          shSend[sid].sign = recvShared[i].sign;      // receiver's own ownership flag
          sid++;
        }
        fill[r] += end - start - 1;
      }
      start = end;
    }
  std::vector<int> shRecvCounts;
  exchange_nodes(comm, shSend, shCounts, sharedNodes, shRecvCounts);
}

void build_csr(dlong nrows, const std::vector<dlong>& counts, std::vector<dlong>& rowStarts) {
  rowStarts.assign((size_t)nrows + 1, 0);
  for (dlong i = 0; i < nrows; ++i) rowStarts[i + 1] = rowStarts[i] + counts[i];
}

// gatherLocal / gatherHalo for Signed, Unsigned and Halo kinds
// (LocalSignedSetup / LocalUnsignedSetup / LocalHaloSetup, ogsSetup.cpp:569-860)
void local_setup(libp_ogs_s& o, const NodeVec& nodes) {
  OgsOperator& L = o.gatherLocal;
  OgsOperator& H = o.gatherHalo;
  L.Ncols = H.Ncols = o.N;
  L.NrowsN = o.NlocalP; L.NrowsT = o.NlocalT;
  H.NrowsN = o.NhaloP;  H.NrowsT = o.NhaloT;
  const bool isHaloKind = (o.kind == LIBP_HALO);
  if (isHaloKind) { L.NrowsN = L.NrowsT = 0; }
  std::vector<dlong> lN((size_t)L.NrowsT, 0), lT((size_t)L.NrowsT, 0), hN((size_t)H.NrowsT, 0), hT((size_t)H.NrowsT, 0);
  auto inN = [&](const Node& n) {
    if (o.kind == LIBP_UNSIGNED) return true;
    if (isHaloKind) return n.sign == 2;
    return n.baseId > 0;
  };
  // row lengths, then columns: both passes run over the nodes in parallel with relaxed atomic cursors, which leaves
  // the entries of a row in arrival order; a per-row sort restores the reference's order (ascending local id,
  // ogsSetup.cpp:637-663) - rows are a handful of entries long
  auto bump = [](dlong& c) { return __atomic_fetch_add(&c, 1, __ATOMIC_RELAXED); };
  const size_t nn = nodes.size();
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nn; ++i) {
    const Node& n = nodes[i];
    if (std::abs(n.sign) == 1) {
      if (isHaloKind) continue;
      if (inN(n)) bump(lN[n.newId]);
      bump(lT[n.newId]);
    } else {
      if (inN(n)) bump(hN[n.newId]);
      bump(hT[n.newId]);
    }
  }
  build_csr(L.NrowsT, lN, L.rowStartsN); build_csr(L.NrowsT, lT, L.rowStartsT);
  build_csr(H.NrowsT, hN, H.rowStartsN); build_csr(H.NrowsT, hT, H.rowStartsT);
  L.colIdsN.resize((size_t)L.nnzN()); L.colIdsT.resize((size_t)L.nnzT());
  H.colIdsN.resize((size_t)H.nnzN()); H.colIdsT.resize((size_t)H.nnzT());
  auto zero = [](std::vector<dlong>& v) {
    const size_t m = v.size();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < m; ++i) v[i] = 0;
  };
  zero(lN); zero(lT); zero(hN); zero(hT);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nn; ++i) {
    const Node& n = nodes[i];
    const dlong g = n.newId;
    if (std::abs(n.sign) == 1) {
      if (isHaloKind) continue;
      if (inN(n)) L.colIdsN[L.rowStartsN[g] + bump(lN[g])] = n.localId;
      L.colIdsT[L.rowStartsT[g] + bump(lT[g])] = n.localId;
    } else {
      if (inN(n)) H.colIdsN[H.rowStartsN[g] + bump(hN[g])] = n.localId;
      H.colIdsT[H.rowStartsT[g] + bump(hT[g])] = n.localId;
    }
  }
  auto sort_rows = [](dlong nrows, const std::vector<dlong>& rs, std::vector<dlong>& cols) {
#pragma omp parallel for schedule(static)
    for (dlong r = 0; r < nrows; ++r)
      if (rs[r + 1] - rs[r] > 1) std::sort(cols.begin() + rs[r], cols.begin() + rs[r + 1]);
  };
  sort_rows(L.NrowsT, L.rowStartsN, L.colIdsN); sort_rows(L.NrowsT, L.rowStartsT, L.colIdsT);
  sort_rows(H.NrowsT, H.rowStartsN, H.colIdsN); sort_rows(H.NrowsT, H.rowStartsT, H.colIdsT);
}

// ---- one rank, Signed / Unsigned kinds: the whole setup on the sorted (key, index) pairs ------------------------
// With a single rank the rendezvous rank is the rank itself, nothing is shared, and FindSharedNodes /
// ConstructSharedNodes / Local*Setup (ogsSetup.cpp:192-331, 333-497, 569-860) look at the same records three times in
// three orders.  Only ONE piece of that depends on an order: the owner of a group is the (rand() % size)-th record in
// the tie order std::sort leaves (the exact sort below).  Everything else is a property of the group (sign, size,
// first local appearance) or of the local order (row numbers by first appearance, columns ascending), so the 32-byte
// records never have to be permuted: one sort of 12-byte pairs, one pass over the groups, prefix sums, one fill.
// `orig[k]` = position of compressed record k in the caller's id array.
void single_rank_setup(libp_ogs_s& o, NodeVec& nodes, const std::vector<dlong, default_init_alloc<dlong>>& orig,
                       hlong* ids) {
  SubLap sub;
  const size_t n = nodes.size();
  std::vector<KeyIdx, default_init_alloc<KeyIdx>> k(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) { k[i].key = (unsigned long long)habs(nodes[i].baseId); k[i].idx = (dlong)i; }
  ExactSort::sort(k.data(), n);
  sub("1 rank: exact sort of (|id|, index)");

  // chunks of whole groups
  const int nch = kSetupChunks;
  std::vector<size_t> cb((size_t)nch + 1), g0((size_t)nch + 1, 0);
  {
    const size_t per = (n + nch - 1) / nch;
    for (int c = 0; c <= nch; ++c) {
      size_t b = std::min(n, (size_t)c * per);
      while (b > 0 && b < n && k[b].key == k[b - 1].key) ++b;
      cb[(size_t)c] = b;
    }
    cb[(size_t)nch] = n;
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nch; ++c) {
      size_t g = 0;
      for (size_t i = cb[(size_t)c]; i < cb[(size_t)c + 1]; ++i) g += (i == cb[(size_t)c]) || k[i].key != k[i - 1].key;
      g0[(size_t)c + 1] = g;
    }
    for (int c = 0; c < nch; ++c) g0[(size_t)c + 1] += g0[(size_t)c];
  }
  const size_t ngroups = g0[(size_t)nch];
  std::vector<int, default_init_alloc<int>> draws;
  if (o.unique) GlibcRandBulk::take(draws, ngroups);
  sub("1 rank: group chunks + rand() draws");

  // group pass: owner, sign, first appearance, sizes; groupStart[g] = first pair of group g
  std::vector<dlong, default_init_alloc<dlong>> firstPos(ngroups), groupStart(ngroups + 1), cntN(ngroups);
  std::vector<signed char, default_init_alloc<signed char>> groupSign(ngroups);
  groupStart[ngroups] = (dlong)n;
  const bool unsignedKind = (o.kind == LIBP_UNSIGNED);
  int is_unique = 1;
  long long nLT = 0, nLP = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(min : is_unique) reduction(+ : nLT, nLP)
  for (int c = 0; c < nch; ++c) {
    size_t g = g0[(size_t)c];
    size_t start = cb[(size_t)c];
    const size_t cend = cb[(size_t)c + 1];
    for (size_t i = start; i < cend; ++i) {
      if (i == cend - 1 || k[i].key != k[i + 1].key) {
        const size_t end = i + 1;
        int positiveCount = 0;
        if (o.unique) {
          const hlong baseId = (hlong)k[start].key;
          const int m = draws[g] % (int)(end - start);
          for (size_t j = start; j < end; ++j) nodes[(size_t)k[j].idx].baseId = -baseId;
          nodes[(size_t)k[start + m].idx].baseId = baseId;
          positiveCount = 1;
        } else {
          for (size_t j = start; j < end; ++j) positiveCount += nodes[(size_t)k[j].idx].baseId > 0;
          if (positiveCount != 1) is_unique = 0;
        }
        const int sign = positiveCount > 0 ? 1 : -1;  // a group without an owner copy only takes part in `Trans` maps
        dlong fp = k[start].idx;
        for (size_t j = start; j < end; ++j) {
          Node& nd = nodes[(size_t)k[j].idx];
          nd.sign = sign;
          nd.newId = (dlong)g;
          fp = std::min(fp, k[j].idx);
        }
        firstPos[g] = fp;
        groupSign[g] = (signed char)sign;
        groupStart[g] = (dlong)start;
        cntN[g] = unsignedKind ? (dlong)(end - start) : (dlong)positiveCount;
        nLT++;
        if (sign == 1) nLP++;
        ++g;
        start = end;
      }
    }
  }
  o.gather_defined = (is_unique == 1);
  o.NlocalT = (dlong)nLT; o.NlocalP = (dlong)nLP; o.NhaloT = o.NhaloP = 0;
  o.Ngather = o.NlocalP;
  o.NgatherGlobal = o.Ngather;
  sub("1 rank: group pass");

  // rows numbered by first appearance in local order, owner rows first (ogsSetup.cpp:411-433)
  std::vector<dlong, default_init_alloc<dlong>> indexMap(ngroups);
  {
    const size_t per = (n + nch - 1) / nch;
    std::vector<std::array<dlong, 2>> cnt((size_t)nch + 1, std::array<dlong, 2>{0, 0});
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nch; ++c) {
      std::array<dlong, 2> q{0, 0};
      const size_t b = std::min(n, (size_t)c * per), e = std::min(n, b + per);
      for (size_t i = b; i < e; ++i) {
        const dlong g = nodes[i].newId;
        if (firstPos[(size_t)g] == (dlong)i) q[groupSign[(size_t)g] == 1 ? 0 : 1]++;
      }
      cnt[(size_t)c + 1] = q;
    }
    cnt[0] = {0, o.NlocalP};
    for (int c = 0; c < nch; ++c) { cnt[(size_t)c + 1][0] += cnt[(size_t)c][0]; cnt[(size_t)c + 1][1] += cnt[(size_t)c][1]; }
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nch; ++c) {
      std::array<dlong, 2> q = cnt[(size_t)c];
      const size_t b = std::min(n, (size_t)c * per), e = std::min(n, b + per);
      for (size_t i = b; i < e; ++i) {
        const dlong g = nodes[i].newId;
        if (firstPos[(size_t)g] == (dlong)i) indexMap[(size_t)g] = q[groupSign[(size_t)g] == 1 ? 0 : 1]++;
      }
    }
  }
  if (o.unique) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) ids[orig[i]] = nodes[i].baseId;
  }
  sub("1 rank: first-appearance renumbering");

  // gatherLocal straight from the groups: row = indexMap[group], columns = the members' positions, ascending
  OgsOperator& L = o.gatherLocal;
  OgsOperator& H = o.gatherHalo;
  L.Ncols = H.Ncols = o.N;
  L.NrowsN = o.NlocalP; L.NrowsT = o.NlocalT;
  H.NrowsN = H.NrowsT = 0;
  H.rowStartsN.assign(1, 0); H.rowStartsT.assign(1, 0);
  H.colIdsN.clear(); H.colIdsT.clear();
  L.rowStartsN.resize((size_t)L.NrowsT + 1); L.rowStartsT.resize((size_t)L.NrowsT + 1);
  L.rowStartsN[0] = L.rowStartsT[0] = 0;
#pragma omp parallel for schedule(static)
  for (size_t g = 0; g < ngroups; ++g) {
    const size_t r = (size_t)indexMap[g];
    L.rowStartsT[r + 1] = groupStart[g + 1] - groupStart[g];
    L.rowStartsN[r + 1] = cntN[g];
  }
  for (size_t r = 0; r < (size_t)L.NrowsT; ++r) {
    L.rowStartsT[r + 1] += L.rowStartsT[r];
    L.rowStartsN[r + 1] += L.rowStartsN[r];
  }
  L.colIdsN.resize((size_t)L.nnzN()); L.colIdsT.resize((size_t)L.nnzT());
#pragma omp parallel for schedule(static)
  for (size_t g = 0; g < ngroups; ++g) {
    const size_t r = (size_t)indexMap[g];
    const size_t b = (size_t)groupStart[g], e = (size_t)groupStart[g + 1];
    dlong* ct = L.colIdsT.data() + L.rowStartsT[r];
    for (size_t j = b; j < e; ++j) ct[j - b] = k[j].idx;  // compressed positions ascend with the caller's positions
    if (e - b > 1) std::sort(ct, ct + (e - b));
    dlong* cn = L.colIdsN.data() + L.rowStartsN[r];
    size_t m = 0;
    for (size_t j = 0; j < e - b; ++j) {
      const dlong c = ct[j];
      if (unsignedKind || nodes[(size_t)c].baseId > 0) cn[m++] = orig[(size_t)c];
    }
    for (size_t j = 0; j < e - b; ++j) ct[j] = orig[(size_t)ct[j]];
  }
  sub("1 rank: gatherLocal maps");
}

// Pairwise exchange lists + post-exchange combine operator (ogsPairwise.cpp:194-415)
void pairwise_setup(libp_ogs_s& o, NodeVec& sharedNodes) {
  const libp_comm_s& comm = *o.comm;
  const int size = comm.size;
  const dlong Nhalo = o.gatherHalo.NrowsT, NhaloP = o.gatherHalo.NrowsN;
  std::sort(sharedNodes.begin(), sharedNodes.end(), [](const Node& a, const Node& b) {
    if (a.rank < b.rank) return true;
    if (a.rank > b.rank) return false;
    return a.newId < b.newId;
  });
  std::vector<int> sendCountsT(size, 0), sendCountsN(size, 0), recvCountsT(size, 0), recvCountsN(size, 0);
  for (const Node& s : sharedNodes) {
    if (s.sign > 0) sendCountsN[s.rank]++;
    sendCountsT[s.rank]++;
  }
  comm.alltoall(sendCountsN.data(), recvCountsN.data(), sizeof(int));
  for (const Node& s : sharedNodes) {
    if (s.sign == 2) o.exN.sendIds.push_back(s.newId);
    o.exT.sendIds.push_back(s.newId);
  }
  NodeVec recvNodes;
  exchange_nodes(comm, sharedNodes, sendCountsT, recvNodes, recvCountsT);
  const dlong Nrecv = (dlong)recvNodes.size();

  OgsOperator& P = o.postmpi;
  P.NrowsN = P.NrowsT = Nhalo;
  P.Ncols = Nhalo + Nrecv;
  std::vector<dlong> cN((size_t)Nhalo, 0), cT((size_t)Nhalo, 1);
  for (dlong n = 0; n < NhaloP; ++n) cN[n] = 1;
  for (const Node& r : recvNodes) {
    if (r.sign == 2) cN[r.localId]++;
    cT[r.localId]++;
  }
  build_csr(Nhalo, cN, P.rowStartsN); build_csr(Nhalo, cT, P.rowStartsT);
  P.colIdsN.assign((size_t)P.nnzN(), 0); P.colIdsT.assign((size_t)P.nnzT(), 0);
  std::fill(cN.begin(), cN.end(), 0); std::fill(cT.begin(), cT.end(), 0);
  for (dlong n = 0; n < NhaloP; ++n) P.colIdsN[P.rowStartsN[n] + cN[n]++] = n;  // own value first
  for (dlong n = 0; n < Nhalo; ++n) P.colIdsT[P.rowStartsT[n] + cT[n]++] = n;
  dlong cnt = Nhalo;
  for (dlong n = 0; n < Nrecv; ++n) {  // then received copies in arrival (= ascending source rank) order
    const dlong id = recvNodes[n].localId;
    if (recvNodes[n].sign == 2) P.colIdsN[P.rowStartsN[id] + cN[id]++] = cnt++;
    P.colIdsT[P.rowStartsT[id] + cT[id]++] = n + Nhalo;
  }

  auto compress = [&](const std::vector<int>& sc, const std::vector<int>& rc, ExchangeLists& ex) {
    int so = 0, ro = 0;
    for (int r = 0; r < size; ++r) {
      if (sc[r] > 0) { ex.sendRanks.push_back(r); ex.sendCounts.push_back(sc[r]); ex.sendOffsets.push_back(so); }
      so += sc[r];
      if (rc[r] > 0) { ex.recvRanks.push_back(r); ex.recvCounts.push_back(rc[r]); ex.recvOffsets.push_back(ro); }
      ro += rc[r];
    }
    ex.sendOffsets.push_back(so);
    ex.recvOffsets.push_back(ro);
  };
  compress(sendCountsN, recvCountsN, o.exN);
  compress(sendCountsT, recvCountsT, o.exT);
}

// Peer-window exchange setup: carve my receive buffers / flags out of my window, tell every neighbour where its
// values land, and precompute the remote address of every value I send (both parities, both flavours).
void p2p_setup(libp_ogs_s& o) {
  libp_comm_s& c = *o.comm;
  P2PExchange& x = o.p2p;
  const int size = c.size, rank = c.rank;
  x.cap = (size_t)std::max<dlong>(o.exT.Nrecv(), 1);
  const size_t recv_off = c.win_alloc(2 * x.cap * sizeof(double));
  const size_t flags_off = c.win_alloc(2 * (size_t)size * sizeof(unsigned long long));
  const size_t acks_off = c.win_alloc(2 * (size_t)size * sizeof(unsigned long long));
  x.recv = reinterpret_cast<double*>(c.win + recv_off);
  x.flags = reinterpret_cast<unsigned long long*>(c.win + flags_off);
  x.acks = reinterpret_cast<unsigned long long*>(c.win + acks_off);
  CUDA_CHECK(cudaMalloc(&x.d_seq, sizeof(unsigned long long)));
  CUDA_CHECK(cudaMemset(x.d_seq, 0, sizeof(unsigned long long)));
  CUDA_CHECK(cudaMalloc(&x.d_done, 2 * sizeof(unsigned int)));
  CUDA_CHECK(cudaMemset(x.d_done, 0, 2 * sizeof(unsigned int)));
  // what each peer needs to know about my window: {recv_off, cap, flags_off, acks_off, slotN, slotT}
  std::vector<int64_t> mine((size_t)size * 6, -1), theirs((size_t)size * 6, -1);
  for (int r = 0; r < size; ++r) {
    int64_t* m = &mine[(size_t)r * 6];
    m[0] = (int64_t)recv_off; m[1] = (int64_t)x.cap; m[2] = (int64_t)flags_off; m[3] = (int64_t)acks_off;
  }
  for (size_t i = 0; i < o.exN.recvRanks.size(); ++i) mine[(size_t)o.exN.recvRanks[i] * 6 + 4] = o.exN.recvOffsets[i];
  for (size_t i = 0; i < o.exT.recvRanks.size(); ++i) mine[(size_t)o.exT.recvRanks[i] * 6 + 5] = o.exT.recvOffsets[i];
  c.alltoall(mine.data(), theirs.data(), 6 * sizeof(int64_t));  // also orders every rank's window zero-fill before use
  for (int f = 0; f < 2; ++f) {
    const ExchangeLists& ex = f == 0 ? o.exN : o.exT;
    std::vector<double*> d0((size_t)ex.Nsend()), d1((size_t)ex.Nsend());
    std::vector<unsigned long long*> pf(ex.sendRanks.size());
    for (size_t i = 0; i < ex.sendRanks.size(); ++i) {
      const int r = ex.sendRanks[i];
      const int64_t* t = &theirs[(size_t)r * 6];
      LIBP_CHECK(t[4 + f] >= 0, "peer-window setup: neighbour lists are not symmetric");
      double* base = reinterpret_cast<double*>(c.peer_win[r] + t[0]) + t[4 + f];
      for (int n = ex.sendOffsets[i]; n < ex.sendOffsets[i + 1]; ++n) {
        d0[n] = base + (n - ex.sendOffsets[i]);
        d1[n] = d0[n] + t[1];
      }
      pf[i] = reinterpret_cast<unsigned long long*>(c.peer_win[r] + t[2]);
    }
    x.sendDst[f][0].upload(d0);
    x.sendDst[f][1].upload(d1);
    x.peerFlags[f].upload(pf);
    x.sendRanks[f].upload(ex.sendRanks);
    x.recvRanks[f].upload(ex.recvRanks);
  }
  std::vector<unsigned long long*> pa(o.exT.recvRanks.size());
  for (size_t i = 0; i < o.exT.recvRanks.size(); ++i)
    pa[i] = reinterpret_cast<unsigned long long*>(c.peer_win[o.exT.recvRanks[i]] + theirs[(size_t)o.exT.recvRanks[i] * 6 + 3]);
  x.peerAcks.upload(pa);
  x.nAckRanks = (int)pa.size();
  (void)rank;
  x.enabled = true;
}

}  // namespace

void libp_b200::OgsOperator::to_device() {
  d_rowStartsN.upload(rowStartsN);
  d_rowStartsT.upload(rowStartsT);
  d_colIdsN.upload(colIdsN);
  d_colIdsT.upload(colIdsT);
}

libp_ogs_s::~libp_ogs_s() {
  if (ev_ready) cudaEventDestroy(ev_ready);
  if (ev_done) cudaEventDestroy(ev_done);
  if (p2p.d_seq) cudaFree(p2p.d_seq);
  if (p2p.d_done) cudaFree(p2p.d_done);
}

void libp_ogs_s::alloc_buffers(size_t bytes_per_node) {
  const size_t needH = (size_t)std::max<dlong>(postmpi.nnzT(), 1) * bytes_per_node;
  const size_t needS = (size_t)std::max<dlong>(exT.Nsend(), 1) * bytes_per_node;
  if (haloBuf.n < needH) haloBuf.alloc(needH);
  if (sendBuf.n < needS) sendBuf.alloc(needS);
}

extern "C" int libp_ogs_setup(libp_dlong N, libp_hlong* ids, libp_comm_t comm, int kind, int unique, int verbose,
                              libp_ogs_t* out) {
  LIBP_API_BEGIN
  (void)verbose;
  LIBP_CHECK(out != nullptr, "null output handle");
  LIBP_CHECK(comm != nullptr, "null communicator");
  LIBP_CHECK(N >= 0 && (N == 0 || ids != nullptr), "bad ids");
  LIBP_CHECK(kind == LIBP_UNSIGNED || kind == LIBP_SIGNED || kind == LIBP_HALO, "bad kind");
  LIBP_CHECK(!((kind == LIBP_UNSIGNED && unique) || (kind == LIBP_HALO && unique)), "Invalid ogs setup requested");
  std::unique_ptr<libp_ogs_s> o(new libp_ogs_s());
  o->comm = comm;
  o->N = N;
  o->kind = kind;
  o->unique = unique != 0;
  const int rank = comm->rank, size = comm->size;

  const bool timing = getenv("LIBP_OGS_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = tnow();
  auto lap = [&](const char* what) {
    if (timing) { const double t1 = tnow(); fprintf(stderr, "[ogs setup] %-28s %.3f s\n", what, t1 - t0); t0 = t1; }
  };
  // compressed list of the non-zero ids (parallel: per-chunk counts, prefix, fill)
  NodeVec nodes;
  std::vector<dlong, default_init_alloc<dlong>> compact((size_t)N);  // position of id n in the compressed list
  std::vector<dlong, default_init_alloc<dlong>> orig;                // ... and back
  {
    const int nchunks = 256;
    std::vector<size_t> cnt((size_t)nchunks + 1, 0);
    const size_t per = ((size_t)N + nchunks - 1) / nchunks;
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nchunks; ++c) {
      size_t k = 0;
      const size_t b = std::min<size_t>((size_t)N, c * per), e = std::min<size_t>((size_t)N, b + per);
      for (size_t n = b; n < e; ++n) k += ids[n] != 0;
      cnt[(size_t)c + 1] = k;
    }
    for (int c = 0; c < nchunks; ++c) cnt[(size_t)c + 1] += cnt[(size_t)c];
    nodes.resize(cnt[(size_t)nchunks]);
    orig.resize(cnt[(size_t)nchunks]);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nchunks; ++c) {
      size_t k = cnt[(size_t)c];
      const size_t b = std::min<size_t>((size_t)N, c * per), e = std::min<size_t>((size_t)N, b + per);
      for (size_t n = b; n < e; ++n) {
        compact[n] = -1;
        if (ids[n] == 0) continue;
        Node nd;
        nd.localId = (dlong)k;
        nd.baseId = (kind == LIBP_UNSIGNED) ? habs(ids[n]) : ids[n];
        nd.newId = 0;
        nd.sign = 0;
        nd.rank = rank;
        nd.destRank = (int)(habs(ids[n]) % size);
        nodes[k] = nd;
        compact[n] = (dlong)k;
        orig[k] = (dlong)n;
        ++k;
      }
    }
  }
  lap("node records");
  NodeVec sharedNodes;
  if (size == 1 && kind != LIBP_HALO && getenv("LIBP_OGS_GENERIC_SETUP") == nullptr) {
    compact.clear();
    compact.shrink_to_fit();
    single_rank_setup(*o, nodes, orig, ids);
    lap("single-rank setup");
  } else {
  find_shared_nodes(*o, nodes);
  lap("find_shared_nodes");
  construct_shared_nodes(*o, nodes, sharedNodes);
  lap("construct_shared_nodes");
  {
    const bool uniq = o->unique;
#pragma omp parallel for schedule(static)
    for (dlong n = 0; n < N; ++n) {
      const dlong c = compact[(size_t)n];
      if (c < 0) continue;
      nodes[(size_t)c].localId = n;
      if (uniq) ids[n] = nodes[(size_t)c].baseId;
    }
    compact.clear();
    compact.shrink_to_fit();
  }
  local_setup(*o, nodes);
  lap("local_setup");
  }
  nodes.clear();
  nodes.shrink_to_fit();
  pairwise_setup(*o, sharedNodes);
  lap("pairwise_setup");

  // device copies; allowed to fail softly when no GPU is present (CPU-side map tests)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
    o->gatherLocal.to_device();
    o->gatherHalo.to_device();
    o->postmpi.to_device();
    o->exN.d_sendIds.upload(o->exN.sendIds);
    o->exT.d_sendIds.upload(o->exT.sendIds);
    CUDA_CHECK(cudaEventCreateWithFlags(&o->ev_ready, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&o->ev_done, cudaEventDisableTiming));
    if (comm->p2p && size > 1) p2p_setup(*o);
  } else {
    cudaGetLastError();
  }
  *out = o.release();
  LIBP_API_END
}

// Development / test hook: the parallel exact sort against std::sort on n records with `nkeys` distinct keys (many
// ties); *same = 1 when the two permutations are identical.
extern "C" int libp_ogs_sort_selftest(libp_dlong n, libp_dlong nkeys, unsigned int seed, int* same) {
  LIBP_API_BEGIN
  LIBP_CHECK(n >= 0 && nkeys >= 1 && same, "bad argument");
  NodeVec a((size_t)n);
  unsigned long long st = seed * 2654435761ull + 12345ull;
  for (dlong i = 0; i < n; ++i) {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    Node nd{};
    nd.localId = i;
    nd.baseId = (hlong)((st >> 33) % (unsigned long long)nkeys) + 1;
    if ((st >> 20) & 1) nd.baseId = -nd.baseId;
    a[(size_t)i] = nd;
  }
  NodeVec b(a);
  std::sort(a.begin(), a.end(), [](const Node& x, const Node& y) { return habs(x.baseId) < habs(y.baseId); });
  exact_sort_nodes(b, [](const Node& x) { return (unsigned long long)habs(x.baseId); });
  int ok = 1;
  for (dlong i = 0; i < n; ++i)
    if (a[(size_t)i].localId != b[(size_t)i].localId) { ok = 0; break; }
  *same = ok;
  LIBP_API_END
}

extern "C" int libp_ogs_rand_selftest(unsigned int seed, libp_dlong n, int* same) {
  LIBP_API_BEGIN
  LIBP_CHECK(n >= 0 && same, "bad argument");
  // n draws in bulk + 100 through rand() must equal n + 100 draws through rand() from the same seed
  std::vector<int> ref((size_t)n + 100);
  srand(seed);
  for (int& v : ref) v = rand();
  srand(seed);
  std::vector<int, default_init_alloc<int>> got;
  GlibcRandBulk::take(got, (size_t)n);
  int ok = 1;
  for (dlong i = 0; i < n; ++i)
    if (got[(size_t)i] != ref[(size_t)i]) { ok = 0; break; }
  for (int i = 0; i < 100; ++i)
    if (rand() != ref[(size_t)n + i]) { ok = 0; break; }
  *same = ok;
  LIBP_API_END
}

extern "C" int libp_ogs_free(libp_ogs_t ogs) {
  LIBP_API_BEGIN
  delete ogs;
  LIBP_API_END
}

extern "C" int libp_ogs_info(libp_ogs_t o, libp_ogs_info_t* info) {
  LIBP_API_BEGIN
  LIBP_CHECK(o && info, "null argument");
  info->N = o->N; info->Ngather = o->Ngather;
  info->NlocalT = o->NlocalT; info->NlocalP = o->NlocalP;
  info->NhaloT = o->NhaloT; info->NhaloP = o->NhaloP;
  info->Nhalo = o->NhaloT - o->NhaloP;
  info->NgatherGlobal = o->NgatherGlobal;
  info->gather_defined = o->gather_defined ? 1 : 0;
  info->NranksSendN = (int)o->exN.sendRanks.size(); info->NranksSendT = (int)o->exT.sendRanks.size();
  info->NranksRecvN = (int)o->exN.recvRanks.size(); info->NranksRecvT = (int)o->exT.recvRanks.size();
  info->NsendN = o->exN.Nsend(); info->NsendT = o->exT.Nsend();
  info->NrecvN = o->exN.Nrecv(); info->NrecvT = o->exT.Nrecv();
  LIBP_API_END
}

extern "C" int libp_ogs_maps(libp_ogs_t o, int which, libp_dlong* NrowsN, libp_dlong* NrowsT,
                             const libp_dlong** rowStartsN, const libp_dlong** rowStartsT,
                             const libp_dlong** colIdsN, const libp_dlong** colIdsT) {
  LIBP_API_BEGIN
  LIBP_CHECK(o, "null handle");
  LIBP_CHECK(which >= 0 && which <= 2, "which must be 0 (local), 1 (halo) or 2 (postmpi)");
  const OgsOperator& op = which == 0 ? o->gatherLocal : which == 1 ? o->gatherHalo : o->postmpi;
  if (NrowsN) *NrowsN = op.NrowsN;
  if (NrowsT) *NrowsT = op.NrowsT;
  if (rowStartsN) *rowStartsN = op.rowStartsN.data();
  if (rowStartsT) *rowStartsT = op.rowStartsT.data();
  if (colIdsN) *colIdsN = op.colIdsN.data();
  if (colIdsT) *colIdsT = op.colIdsT.data();
  LIBP_API_END
}

extern "C" int libp_ogs_exchange_lists(libp_ogs_t o, int trans, libp_dlong* Nsend, const libp_dlong** sendIds,
                                       int* NranksSend, const int** sendRanks, const int** sendCounts,
                                       const int** sendOffsets, int* NranksRecv, const int** recvRanks,
                                       const int** recvCounts, const int** recvOffsets) {
  LIBP_API_BEGIN
  LIBP_CHECK(o, "null handle");
  const ExchangeLists& ex = (trans == LIBP_NOTRANS) ? o->exN : o->exT;
  if (Nsend) *Nsend = ex.Nsend();
  if (sendIds) *sendIds = ex.sendIds.data();
  if (NranksSend) *NranksSend = (int)ex.sendRanks.size();
  if (sendRanks) *sendRanks = ex.sendRanks.data();
  if (sendCounts) *sendCounts = ex.sendCounts.data();
  if (sendOffsets) *sendOffsets = ex.sendOffsets.data();
  if (NranksRecv) *NranksRecv = (int)ex.recvRanks.size();
  if (recvRanks) *recvRanks = ex.recvRanks.data();
  if (recvCounts) *recvCounts = ex.recvCounts.data();
  if (recvOffsets) *recvOffsets = ex.recvOffsets.data();
  LIBP_API_END
}

extern "C" int libp_ogs_global_to_local(libp_ogs_t o, libp_dlong* g2l) {
  LIBP_API_BEGIN
  LIBP_CHECK(o && g2l, "null argument");
  LIBP_CHECK(o->NgatherGlobal != 0, "ogs handle is not set up.");
  const dlong N = o->N;
#pragma omp parallel for schedule(static)
  for (dlong n = 0; n < N; ++n) g2l[n] = -1;
  const OgsOperator* ops[2] = {&o->gatherLocal, &o->gatherHalo};
  const dlong offs[2] = {0, o->NlocalT};
  for (int w = 0; w < 2; ++w) {
    const OgsOperator& op = *ops[w];
    const dlong off = offs[w];
#pragma omp parallel for schedule(static)
    for (dlong r = 0; r < op.NrowsT; ++r)  // every local node sits in exactly one row: no write conflicts
      for (dlong g = op.rowStartsT[r]; g < op.rowStartsT[r + 1]; ++g) g2l[op.colIdsT[g]] = r + off;
  }
  LIBP_API_END
}
