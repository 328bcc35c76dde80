// sparse_direct.cpp — host analysis of the sparse direct step solve of one large system (large.cu).
//
// Replaces, once per topology, what faer's SymbolicLlt::try_new does inside precompute_symbolic_cholesky
// (ezpz/src/solver.rs:289-300: fill-reducing ordering, elimination tree, symbolic factorisation) with a
// schedule a GPU can run: the columns of L grouped by their height in the elimination tree, so that all
// columns of one level factorise in parallel.
//
//   1. adjacency of the graph of A = JtJ + lambda*I (from lower(A), structure.cpp);
//   2. elimination order: natural (0..n-1, which keeps the arithmetic identical to the small-system tape and
//      to the oracle's default) when the natural elimination tree is shallow — block-diagonal systems such as
//      massive_parallel_system — otherwise nested dissection on breadth-first level structures: one BFS per
//      connected component from a pseudo-peripheral vertex, separators = whole BFS level sets chosen by
//      recursive bisection of the level range, leaves in (level, index) order.  Sketch graphs are chains and
//      trees of small cells, so level sets are a cell wide (about 13 variables on the 1M-variable sketch of
//      BASELINE.json config 4) and the tree is O(log n) separators deep;
//   3. symbolic Cholesky in that order (column merge over the elimination tree), entries flagged when A itself
//      has them (pure fill starts from +0.0 and skips the JtJ products);
//   4. L by rows (value order) and by columns in level order (work lists), the level pointer table, and the
//      first level from which a single CTA runs the rest of the tree (the top of the tree is a few columns
//      per level: one __syncthreads per level instead of a grid barrier).
//
// Arithmetic-order spec in elimination numbering j <-> variable perm[j]: identical formulas to
// the natural-order oracle applied to P A Pt; tests pass `perm` to the oracle to check bit-exactness.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <vector>

#include "host_parallel.h"
#include "structure.h"

namespace ezs {

namespace {

constexpr uint32_t kNone = UINT32_MAX;
constexpr uint32_t kNaturalMaxHeight = 192;     // natural order is kept when its elimination tree is this shallow
constexpr uint32_t kLeafVars = 48;              // dissection stops at parts of this many variables
constexpr uint64_t kMaxFactorEntries = 1ull << 29;  // beyond this nnz(L) the PCG path is used
constexpr uint32_t kMaxPanelWidth = 16;         // columns per supernode panel
constexpr uint32_t kWarpPanelCap = 576;         // doubles of panel a warp team can stage (large.cu: TeamCaps<32>::panel)
constexpr uint32_t kLanePanel = 8;              // panels of at most this many doubles are factorised by a single thread

struct Graph {
    uvec<uint32_t> ptr, adj;  // symmetric adjacency without self loops
};

Graph build_graph(const ezpz_structure& S) {
    // bucket pass over ranges of columns of lower(A) on host threads: count both ends of every off-diagonal entry, scatter,
    // sort every (short) neighbour list
    const uint32_t n = S.n;
    const bool shared = host_threads(n, kHostGrain) > 1;
    Graph g;
    g.ptr.resize((size_t)n + 2);
    parallel_fill(g.ptr.data(), g.ptr.size(), 0u);
    parallel_ranges(n, kHostGrain, [&](uint32_t jb, uint32_t je, uint32_t) {
        for (uint32_t j = jb; j < je; ++j)
            for (uint32_t p = S.a_col_ptr[j]; p < S.a_col_ptr[j + 1]; ++p) {
                const uint32_t i = S.a_row_idx[p];
                if (i == j) continue;
                bump(&g.ptr[i + 2], shared);
                bump(&g.ptr[j + 2], shared);
            }
    });
    for (uint32_t v = 0; v < n; ++v) g.ptr[v + 2] += g.ptr[v + 1];  // g.ptr[v + 1] = cursor of v's list
    g.adj.resize(g.ptr[(size_t)n + 1]);
    parallel_ranges(n, kHostGrain, [&](uint32_t jb, uint32_t je, uint32_t) {
        for (uint32_t j = jb; j < je; ++j)
            for (uint32_t p = S.a_col_ptr[j]; p < S.a_col_ptr[j + 1]; ++p) {
                const uint32_t i = S.a_row_idx[p];
                if (i == j) continue;
                g.adj[bump(&g.ptr[i + 1], shared)] = j;
                g.adj[bump(&g.ptr[j + 1], shared)] = i;
            }
    });
    g.ptr.pop_back();  // (the cursors have advanced by one vertex: ptr[v] .. ptr[v + 1] is v's list)
    parallel_ranges(n, kHostGrain, [&](uint32_t vb, uint32_t ve, uint32_t) {
        for (uint32_t v = vb; v < ve; ++v)
            if (!std::is_sorted(g.adj.begin() + g.ptr[v], g.adj.begin() + g.ptr[v + 1])) std::sort(g.adj.begin() + g.ptr[v], g.adj.begin() + g.ptr[v + 1]);
    });
    return g;
}

// Elimination tree of P A Pt (Liu's algorithm with path compression) and the height of every node's subtree
// (leaves = 0).  Returns the number of levels.
uint32_t etree_levels(const Graph& g, const std::vector<uint32_t>& perm, const std::vector<uint32_t>& iperm,
                      std::vector<uint32_t>& parent, std::vector<uint32_t>& level) {
    const uint32_t n = (uint32_t)perm.size();
    parent.assign(n, kNone);
    std::vector<uint32_t> anc(n, kNone);
    for (uint32_t j = 0; j < n; ++j) {
        const uint32_t v = perm[j];
        for (uint32_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
            uint32_t r = iperm[g.adj[p]];
            if (r >= j) continue;
            while (anc[r] != kNone && anc[r] != j) {
                const uint32_t next = anc[r];
                anc[r] = j;
                r = next;
            }
            if (anc[r] == kNone) {
                anc[r] = j;
                parent[r] = j;
            }
        }
    }
    level.assign(n, 0);
    uint32_t height = 0;
    for (uint32_t j = 0; j < n; ++j) {
        height = std::max(height, level[j] + 1);
        if (parent[j] != kNone) level[parent[j]] = std::max(level[parent[j]], level[j] + 1);
    }
    return n ? height : 0;
}

// Nested dissection on BFS level structures (see the header comment).
void nested_dissection(const Graph& g, uint32_t n, std::vector<uint32_t>& perm) {
    perm.clear();
    perm.reserve(n);
    std::vector<uint32_t> stamp(n, 0), lev(n, 0), queue, comp, lstart;
    uint32_t cur_stamp = 0;
    std::vector<uint8_t> done(n, 0);
    auto bfs = [&](uint32_t start, std::vector<uint32_t>& out) {
        ++cur_stamp;
        out.clear();
        out.push_back(start);
        stamp[start] = cur_stamp;
        lev[start] = 0;
        for (size_t h = 0; h < out.size(); ++h) {
            const uint32_t v = out[h];
            for (uint32_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                const uint32_t u = g.adj[p];
                if (stamp[u] != cur_stamp) {
                    stamp[u] = cur_stamp;
                    lev[u] = lev[v] + 1;
                    out.push_back(u);
                }
            }
        }
    };
    struct Range {
        uint32_t a, b;
        bool emit_sep;
        uint32_t sep;
    };
    std::vector<Range> stack;
    for (uint32_t v0 = 0; v0 < n; ++v0) {
        if (done[v0]) continue;
        bfs(v0, comp);
        if (comp.size() > kLeafVars) {
            // pseudo-peripheral start: restart from a vertex of the last level (smallest degree), twice
            for (int pass = 0; pass < 2; ++pass) {
                const uint32_t last_level = lev[comp.back()];
                uint32_t best = comp.back(), best_deg = kNone;
                for (size_t q = comp.size(); q-- > 0 && lev[comp[q]] == last_level;) {
                    const uint32_t d = g.ptr[comp[q] + 1] - g.ptr[comp[q]];
                    if (d < best_deg || (d == best_deg && comp[q] < best)) {
                        best_deg = d;
                        best = comp[q];
                    }
                }
                const uint32_t old_depth = last_level;
                bfs(best, comp);
                if (lev[comp.back()] <= old_depth) break;
            }
        }
        for (uint32_t v : comp) done[v] = 1;
        // level structure: comp is already in BFS order = non-decreasing level; sort each level by index
        const uint32_t n_lev = lev[comp.back()] + 1;
        lstart.assign((size_t)n_lev + 1, 0);
        for (uint32_t v : comp) lstart[lev[v] + 1]++;
        for (uint32_t l = 0; l < n_lev; ++l) lstart[l + 1] += lstart[l];
        for (uint32_t l = 0; l < n_lev; ++l) std::sort(comp.begin() + lstart[l], comp.begin() + lstart[l + 1]);
        // recursive bisection of the level range [0, n_lev), separators last
        stack.clear();
        stack.push_back({0, n_lev, false, 0});
        while (!stack.empty()) {
            const Range r = stack.back();
            stack.pop_back();
            if (r.emit_sep) {
                perm.insert(perm.end(), comp.begin() + lstart[r.sep], comp.begin() + lstart[r.sep + 1]);
                continue;
            }
            const uint32_t count = lstart[r.b] - lstart[r.a];
            if (count <= kLeafVars || r.b - r.a < 3) {
                perm.insert(perm.end(), comp.begin() + lstart[r.a], comp.begin() + lstart[r.b]);
                continue;
            }
            // separator level m in [a+1, b-2] that balances the two sides
            uint32_t best_m = r.a + 1;
            uint64_t best_cost = UINT64_MAX;
            {
                // the balance is monotone in m: binary search the crossing, then look at its neighbours
                uint32_t lo = r.a + 1, hi = r.b - 2;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) / 2;
                    const uint32_t left = lstart[mid] - lstart[r.a], right = lstart[r.b] - lstart[mid + 1];
                    if (left < right) lo = mid + 1;
                    else hi = mid;
                }
                for (uint32_t m = (lo > r.a + 1 ? lo - 1 : lo); m <= std::min(lo + 1, r.b - 2); ++m) {
                    const uint64_t left = lstart[m] - lstart[r.a], right = lstart[r.b] - lstart[m + 1];
                    const uint64_t sep = lstart[m + 1] - lstart[m];
                    const uint64_t cost = (left > right ? left - right : right - left) + 2 * sep;
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_m = m;
                    }
                }
            }
            // processed in reverse push order: left part, right part, then the separator
            stack.push_back({0, 0, true, best_m});
            stack.push_back({best_m + 1, r.b, false, 0});
            stack.push_back({r.a, best_m, false, 0});
        }
    }
}

// The elimination tree cut into independent pieces for the host threads: owner[j] = the thread that owns column j's
// subtree, or n_threads for the columns above the cut.  The largest subtree is split (its root goes above the cut, its
// children become subtrees of their own) until there are several subtrees per thread and none holds more than a small share
// of the columns; subtrees are dealt to the threads largest first, each to the least loaded thread.
struct TreeTasks {
    uint32_t n_threads = 1;
    std::vector<uint8_t> owner;
};

TreeTasks cut_tree(const std::vector<uint32_t>& parent, uint32_t n_threads) {
    const uint32_t n = (uint32_t)parent.size();
    TreeTasks T;
    T.n_threads = std::max(1u, std::min(n_threads, 64u));
    if (T.n_threads == 1) {
        T.owner.assign(n, 0);
        return T;
    }
    std::vector<uint32_t> size(n, 1), child_head(n, kNone), child_next(n, kNone);
    for (uint32_t j = 0; j < n; ++j)
        if (parent[j] != kNone) {
            size[parent[j]] += size[j];
            child_next[j] = child_head[parent[j]];
            child_head[parent[j]] = j;
        }
    std::vector<std::pair<uint32_t, uint32_t>> heap;  // (columns, root)
    for (uint32_t j = 0; j < n; ++j)
        if (parent[j] == kNone) heap.push_back({size[j], j});
    std::make_heap(heap.begin(), heap.end());
    constexpr uint8_t kUnset = 255;
    T.owner.assign(n, kUnset);
    const uint32_t target = 8 * T.n_threads, small = std::max(1024u, n / (8 * T.n_threads));
    uint32_t splits = 0;
    while (!heap.empty() && heap.front().first > small && (heap.size() < target || heap.front().first > n / T.n_threads) &&
           splits < (1u << 16)) {
        std::pop_heap(heap.begin(), heap.end());
        const uint32_t r = heap.back().second;
        heap.pop_back();
        T.owner[r] = (uint8_t)T.n_threads;  // above the cut
        ++splits;
        for (uint32_t c = child_head[r]; c != kNone; c = child_next[c]) {
            heap.push_back({size[c], c});
            std::push_heap(heap.begin(), heap.end());
        }
    }
    std::sort(heap.begin(), heap.end(), [](const auto& x, const auto& y) { return x.first != y.first ? x.first > y.first : x.second < y.second; });
    std::vector<uint64_t> load(T.n_threads, 0);
    for (const auto& task : heap) {
        const uint32_t t = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
        load[t] += task.first;
        T.owner[task.second] = (uint8_t)t;
    }
    for (uint32_t j = n; j-- > 0;)  // parents first: a column below a subtree's root belongs to that root's thread
        if (T.owner[j] == kUnset) T.owner[j] = T.owner[parent[j]];
    return T;
}

// aprod_ptr / aprod_a / aprod_b from the entries A has (aent_colptr, aent_row; elimination numbering, `perm` = position ->
// variable): for every entry the rows its two columns of J share, ascending, as pairs of CSC positions.  Count, prefix, fill;
// ranges of columns on host threads.
void fill_products(ezpz_structure& S, const uint32_t* perm) {
    LargeProgram& P = S.large;
    const uint32_t n = S.n;
    const size_t n_ent = P.aent_row.size();
    auto shared_rows = [&](uint32_t ci, uint32_t cj, auto&& fn) {  // fn(position in ci, position in cj)
        uint32_t pi = S.csc_col_ptr[ci], pj = S.csc_col_ptr[cj];
        const uint32_t pie = S.csc_col_ptr[ci + 1], pje = S.csc_col_ptr[cj + 1];
        while (pi < pie && pj < pje) {
            const uint32_t ri = S.csc_row_idx[pi], rj = S.csc_row_idx[pj];
            if (ri == rj) fn(pi++, pj++);
            else if (ri < rj) ++pi;
            else ++pj;
        }
    };
    P.aprod_ptr.resize(n_ent + 1);
    P.aprod_ptr[0] = 0;
    parallel_ranges(n, kHostGrain, [&](uint32_t c0, uint32_t c1, uint32_t) {
        for (uint32_t j = c0; j < c1; ++j)
            for (uint32_t e = P.aent_colptr[j]; e < P.aent_colptr[j + 1]; ++e) {
                uint32_t found = 0;
                shared_rows(perm[P.aent_row[e]], perm[j], [&](uint32_t, uint32_t) { ++found; });
                P.aprod_ptr[e + 1] = found;
            }
    });
    prefix_sum(P.aprod_ptr);
    const size_t n_prod = P.aprod_ptr[n_ent];
    P.aprod_a.resize(n_prod);
    P.aprod_b.resize(n_prod);
    parallel_ranges(n, kHostGrain, [&](uint32_t c0, uint32_t c1, uint32_t) {
        for (uint32_t j = c0; j < c1; ++j)
            for (uint32_t e = P.aent_colptr[j]; e < P.aent_colptr[j + 1]; ++e) {
                uint32_t at = P.aprod_ptr[e];
                shared_rows(perm[P.aent_row[e]], perm[j], [&](uint32_t pi, uint32_t pj) {
                    P.aprod_a[at] = pi;
                    P.aprod_b[at++] = pj;
                });
            }
    });
}

}  // namespace

// Fills the sparse-direct part of S.large.  Leaves P.direct false when the factor would be too large.
void build_sparse_direct(ezpz_structure& S, const uint32_t* order_hint, bool hint_nested, const ezpz_structure* same_a) {
    LargeProgram& P = S.large;
    const uint32_t n = S.n;
    P.direct = false;
    if (n == 0) return;
    if (same_a && same_a->large.direct && same_a->n == n) {
        // Constraints were added that couple no new pair of variables (ezpz_b200_structure_extend): A has the pattern it had,
        // so the order, the symbolic factorisation, the supernodes, their updates and stages are the base's; only the
        // products behind the entries of A = JtJ run over new rows.
        const LargeProgram& B = same_a->large;
        P.nested = B.nested;
        P.n_levels = B.n_levels;
        P.nnz_l = B.nnz_l;
        const LargeProgram::u32v* from[] = {&B.perm, &B.sn_ptr, &B.sn_row_ptr, &B.sn_rows, &B.panel_off, &B.upd_ptr, &B.upd_sn, &B.upd_rbegin,
                                            &B.upd_ncols, &B.upd_rel_ptr, &B.upd_rel, &B.upd_rec, &B.stage_ptr, &B.stage_sn, &B.stage_rec,
                                            &B.aent_slot, &B.aent_row, &B.aent_colptr, &B.diag_slot};
        LargeProgram::u32v* to[] = {&P.perm, &P.sn_ptr, &P.sn_row_ptr, &P.sn_rows, &P.panel_off, &P.upd_ptr, &P.upd_sn, &P.upd_rbegin,
                                    &P.upd_ncols, &P.upd_rel_ptr, &P.upd_rel, &P.upd_rec, &P.stage_ptr, &P.stage_sn, &P.stage_rec,
                                    &P.aent_slot, &P.aent_row, &P.aent_colptr, &P.diag_slot};
        for (size_t k = 0; k < sizeof(from) / sizeof(from[0]); ++k) {
            to[k]->resize(from[k]->size());
            parallel_copy(to[k]->data(), from[k]->data(), from[k]->size());
        }
        fill_products(S, P.perm.data());
        P.direct = true;
        return;
    }
    const char* force = std::getenv("EZPZ_B200_FORCE_PCG");
    if (force && force[0] == '1') return;
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&, last = t_start](const char* what) mutable {
        const char* dbg = std::getenv("EZPZ_B200_DEBUG");
        if (!(dbg && dbg[0] == '1')) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[sparse_direct] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    };
    const Graph g = build_graph(S);
    lap("adjacency");

    // ---- 2. elimination order ---------------------------------------------------------------------
    std::vector<uint32_t> perm(n), iperm(n), parent, level;
    std::iota(perm.begin(), perm.end(), 0u);
    std::iota(iperm.begin(), iperm.end(), 0u);
    uint32_t n_levels = 0;
    P.nested = false;
    const char* force_nd = std::getenv("EZPZ_B200_FORCE_ND");
    const bool forced = force_nd && force_nd[0] == '1';
    auto dissect = [&](std::vector<uint32_t>& nd, std::vector<uint32_t>& ind, std::vector<uint32_t>& nparent, std::vector<uint32_t>& nlevel) {
        nested_dissection(g, n, nd);
        ind.resize(n);
        for (uint32_t j = 0; j < n; ++j) ind[nd[j]] = j;
        return etree_levels(g, nd, ind, nparent, nlevel);
    };
    auto adopt = [&](std::vector<uint32_t>& nd, std::vector<uint32_t>& ind, std::vector<uint32_t>& nparent, std::vector<uint32_t>& nlevel, uint32_t h) {
        if (h < n_levels || forced) {
            perm.swap(nd);
            iperm.swap(ind);
            parent.swap(nparent);
            level.swap(nlevel);
            n_levels = h;
            P.nested = true;
        }
    };
    std::vector<uint32_t> nd, ind, nparent, nlevel;
    if (order_hint) {
        // an order to keep (constraints were added to an analysed structure: new entries of A, the same variables; any order
        // of the variables is a valid elimination order and the dissection of the old graph still separates the new one as
        // long as the additions are local)
        perm.assign(order_hint, order_hint + n);
        for (uint32_t j = 0; j < n; ++j) iperm[perm[j]] = j;
        n_levels = etree_levels(g, perm, iperm, parent, level);
        P.nested = hint_nested;
    } else if (host_threads(n, 2 * kHostGrain) > 1) {
        // large systems: the natural-order tree (almost always too deep to keep) on a second thread while this one dissects
        std::thread natural([&] { n_levels = etree_levels(g, perm, iperm, parent, level); });
        const uint32_t h = dissect(nd, ind, nparent, nlevel);
        natural.join();
        if (n_levels > kNaturalMaxHeight || forced) adopt(nd, ind, nparent, nlevel, h);
    } else {
        n_levels = etree_levels(g, perm, iperm, parent, level);
        if (n_levels > kNaturalMaxHeight || forced) adopt(nd, ind, nparent, nlevel, dissect(nd, ind, nparent, nlevel));
    }

    lap("ordering");
    // ---- 3. symbolic factorisation, by columns -----------------------------------------------------
    // struct(L_j) = {i > j : A_perm(i, j) != 0}  U  (struct(L_c) \ {j}) over the etree children c of j.
    // A column needs its children only, so disjoint subtrees of the elimination tree are independent: the tree is cut below
    // its top (TreeTasks), every host thread merges the columns of its subtrees in ascending order into a store of its own,
    // the columns above the cut follow on the calling thread, and the stores are copied into column order at the end.
    // An entry is kept as row * 2 + (A has it): one 32-bit sort key per entry.
    uvec<uint32_t> lc_ptr((size_t)n + 1), lc_row;
    uvec<uint8_t> lc_in_a;
    if (n >= (1u << 31)) return;
    {
        const TreeTasks tasks = cut_tree(parent, host_threads(n, kHostGrain));
        const uint32_t nt = tasks.n_threads;
        std::vector<uvec<uint32_t>> store(nt + 1);  // [nt] = the columns above the cut
        uvec<size_t> col_off(n);                    // offset of column j in its owner's store
        std::vector<uint32_t> child_head(n, kNone), child_next(n, kNone);
        for (uint32_t j = n; j-- > 0;)  // (descending: every list ends up ascending)
            if (parent[j] != kNone) {
                child_next[j] = child_head[parent[j]];
                child_head[parent[j]] = j;
            }
        std::vector<uint8_t> too_large(nt + 1, 0);
        auto run = [&](uint32_t t) {
            uvec<uint32_t>& out = store[t];
            std::vector<uint32_t> mark(n, kNone);
            uvec<uint32_t> base, extra;
            out.reserve(t == nt ? 1024 : (size_t)g.adj.size() * 2 / nt);
            for (uint32_t j = 0; j < n; ++j) {
                if (tasks.owner[j] != t) continue;
                const size_t start = out.size();
                col_off[j] = start;
                mark[j] = j;
                // The tallest child's column is sorted already: what it contributes stays in order (base), everything else
                // (A's own entries, what the other children add) is collected, sorted and merged into it.
                extra.clear();
                const uint32_t v = perm[j];
                for (uint32_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                    const uint32_t i = iperm[g.adj[p]];
                    if (i > j && mark[i] != j) {
                        mark[i] = j;
                        extra.push_back(i * 2 + 1);
                    }
                }
                uint32_t tallest = kNone;
                for (uint32_t c = child_head[j]; c != kNone; c = child_next[c])
                    if (tallest == kNone || lc_ptr[c + 1] > lc_ptr[tallest + 1]) tallest = c;
                base.clear();
                auto take = [&](uint32_t c, uvec<uint32_t>& into) {
                    const uvec<uint32_t>& from = store[tasks.owner[c]];  // this thread's own store below the cut
                    const uint32_t* f = from.data() + col_off[c];
                    for (uint32_t p = 0, e = lc_ptr[c + 1]; p < e; ++p) {
                        const uint32_t i = f[p] >> 1;
                        if (mark[i] != j) {  // i == j is marked already
                            mark[i] = j;
                            into.push_back(i * 2);
                        }
                    }
                };
                if (tallest != kNone) take(tallest, base);
                for (uint32_t c = child_head[j]; c != kNone; c = child_next[c])
                    if (c != tallest) take(c, extra);
                std::sort(extra.begin(), extra.end());
                out.resize(start + base.size() + extra.size());
                std::merge(base.begin(), base.end(), extra.begin(), extra.end(), out.begin() + start);
                lc_ptr[j + 1] = (uint32_t)(out.size() - start);  // (lengths; offsets after the prefix sum below)
                if (out.size() > kMaxFactorEntries) {
                    too_large[t] = 1;
                    return;
                }
            }
        };
        HostPool::get().run(nt, run);  // (nt == 1: the caller alone)
        for (uint32_t t = 0; t < nt; ++t)
            if (too_large[t]) return;
        if (nt > 1) run(nt);  // (a single thread owns every column)
        if (too_large[nt]) return;
        lc_ptr[0] = 0;
        uint64_t total = 0;
        for (uint32_t j = 0; j < n; ++j) {
            total += lc_ptr[j + 1];
            if (total > kMaxFactorEntries) return;
            lc_ptr[j + 1] = (uint32_t)total;
        }
        lc_row.resize(total);
        lc_in_a.resize(total);
        parallel_ranges(n, kHostGrain, [&](uint32_t jb, uint32_t je, uint32_t) {
            for (uint32_t j = jb; j < je; ++j) {
                const uint32_t* from = store[tasks.owner[j]].data() + col_off[j];
                for (uint32_t q = lc_ptr[j], k = 0; q < lc_ptr[j + 1]; ++q, ++k) {
                    lc_row[q] = from[k] >> 1;
                    lc_in_a[q] = (uint8_t)(from[k] & 1u);
                }
            }
        });
    }
    const size_t nnz_l_true = lc_row.size();

    lap("symbolic factorisation");
    // ---- 4. supernodes: maximal chains of the elimination tree, at most kMaxPanelWidth columns -------
    // Column j + 1 continues column j's supernode when j's parent is j + 1 and j + 1 has no other child: along such
    // a chain struct(L_{j+1}) contains struct(L_j) \ {j + 1}, so the columns share (almost) one row structure and
    // are factorised as ONE dense panel: rows = the supernode's own columns followed by the union of the
    // sub-diagonal structures (entries the true factor does not have are explicit zeros; they stay exactly zero
    // and contribute fma(-0, x, acc) = acc, so the values of the true entries do not change).
    std::vector<uint32_t> n_child(n, 0);
    for (uint32_t j = 0; j < n; ++j)
        if (parent[j] != kNone) n_child[parent[j]]++;
    // Relaxed: column j also continues the supernode of its child j - 1 when it has further children in other
    // subtrees (their rows become explicit zeros in the earlier columns of the panel): fewer, wider panels and a
    // lower supernode tree.  EZPZ_B200_STRICT_SUPERNODES=1 keeps the strict chains.
    const char* strict = std::getenv("EZPZ_B200_STRICT_SUPERNODES");
    const bool relax = !(strict && strict[0] == '1');
    P.sn_ptr.assign(1, 0u);
    for (uint32_t j = 1; j < n; ++j) {
        const bool chain = parent[j - 1] == j && (n_child[j] == 1 || relax) && j - P.sn_ptr.back() < kMaxPanelWidth;
        if (!chain) P.sn_ptr.push_back(j);
    }
    P.sn_ptr.push_back(n);
    const uint32_t n_sn = (uint32_t)P.sn_ptr.size() - 1;
    std::vector<uint32_t> sn_of(n);
    for (uint32_t s = 0; s < n_sn; ++s)
        for (uint32_t j = P.sn_ptr[s]; j < P.sn_ptr[s + 1]; ++j) sn_of[j] = s;
    // panel rows: own columns, then the sorted union of the columns' sub-diagonal rows outside the supernode.  Supernodes
    // are independent: ranges of them on host threads, each into a list of its own, copied into place afterwards.
    P.sn_row_ptr.resize((size_t)n_sn + 1);
    P.sn_row_ptr[0] = 0;
    P.panel_off.resize((size_t)n_sn + 1);
    {
        struct Part {
            uvec<uint32_t> rows;
            uint32_t s0 = 0;
        };
        std::vector<Part> parts(64);
        uint32_t n_parts = 1;
        parallel_ranges(n_sn, kHostGrain / 4, [&](uint32_t sb, uint32_t se, uint32_t t) {
            Part& part = parts[t];
            part.s0 = sb;
            std::vector<uint32_t> mark(n, kNone);
            uvec<uint32_t> below;
            part.rows.reserve((size_t)(lc_ptr[P.sn_ptr[se]] - lc_ptr[P.sn_ptr[sb]]) / 2 + 1024);
            for (uint32_t s = sb; s < se; ++s) {
                const uint32_t j0 = P.sn_ptr[s], j1 = P.sn_ptr[s + 1];
                below.clear();
                for (uint32_t j = j0; j < j1; ++j)
                    for (uint32_t q = lc_ptr[j]; q < lc_ptr[j + 1]; ++q) {
                        const uint32_t i = lc_row[q];
                        if (i >= j1 && mark[i] != s) {
                            mark[i] = s;
                            below.push_back(i);
                        }
                    }
                // (a one-column supernode's rows are its column: sorted already)
                if (j1 - j0 > 1) std::sort(below.begin(), below.end());
                for (uint32_t j = j0; j < j1; ++j) part.rows.push_back(j);
                part.rows.insert(part.rows.end(), below.begin(), below.end());
                P.sn_row_ptr[s + 1] = (j1 - j0) + (uint32_t)below.size();  // (heights; offsets after the prefix sum)
            }
        }, &n_parts);
        uint64_t off = 0, rows_total = 0;
        for (uint32_t s = 0; s < n_sn; ++s) {
            const uint64_t h = P.sn_row_ptr[s + 1], w = P.sn_ptr[s + 1] - P.sn_ptr[s];
            P.panel_off[s] = (uint32_t)off;
            off += w * h;
            rows_total += h;
            if (off > kMaxFactorEntries || rows_total > UINT32_MAX) return;  // leave P.direct false: PCG path
            P.sn_row_ptr[s + 1] = (uint32_t)rows_total;
        }
        P.panel_off[n_sn] = (uint32_t)off;
        P.sn_rows.resize(rows_total);
        parallel_ranges(n_parts, 1, [&](uint32_t tb, uint32_t te, uint32_t) {
            for (uint32_t t = tb; t < te; ++t)
                std::copy(parts[t].rows.begin(), parts[t].rows.end(), P.sn_rows.begin() + P.sn_row_ptr[parts[t].s0]);
        });
    }
    const uint32_t nnz_l = P.panel_off[n_sn];  // doubles of panel storage (explicit zeros included)
    auto panel_h = [&](uint32_t s) { return P.sn_row_ptr[s + 1] - P.sn_row_ptr[s]; };
    auto panel_w = [&](uint32_t s) { return P.sn_ptr[s + 1] - P.sn_ptr[s]; };
    bool inconsistent = false;  // a row that should be in a panel is not: never expected; falls back to the PCG path
    lap("supernodes and panels");
    // ---- 5. update lists: which descendant supernodes K update supernode J, and where K's rows land in J ----
    // K's rows below its own columns, grouped by the supernode that owns them: every group is one (J <- K) update;
    // the rows of K from the group's start to the end of K's list all lie inside J's panel (elimination tree).
    // Bucket pass on host threads (like the pattern of J): count the updates every J receives, scatter (K, first row, rows)
    // into J's bucket, sort each bucket by K; the relative positions are then filled update by update, K's rows and J's
    // rows walked together (both ascending).
    {
        auto groups_of = [&](uint32_t K, auto&& fn) {  // fn(J, first row of the group in K's list, rows in the group)
            const uint32_t rb = P.sn_row_ptr[K], w = panel_w(K), h = panel_h(K);
            uint32_t t = w;
            while (t < h) {
                const uint32_t J = sn_of[P.sn_rows[rb + t]];
                uint32_t t2 = t;
                while (t2 < h && sn_of[P.sn_rows[rb + t2]] == J) ++t2;
                fn(J, t, t2 - t);
                t = t2;
            }
        };
        const bool shared = host_threads(n_sn, kHostGrain / 4) > 1;
        uvec<uint32_t> cursor((size_t)n_sn + 2);
        parallel_fill(cursor.data(), cursor.size(), 0u);
        parallel_ranges(n_sn, kHostGrain / 4, [&](uint32_t kb, uint32_t ke, uint32_t) {
            for (uint32_t K = kb; K < ke; ++K)
                groups_of(K, [&](uint32_t J, uint32_t, uint32_t) { bump(&cursor[J + 2], shared); });
        });
        for (uint32_t J = 0; J < n_sn; ++J) cursor[J + 2] += cursor[J + 1];  // cursor[J + 1] = start of J's bucket
        const uint32_t n_upd = cursor[(size_t)n_sn + 1];
        struct Upd {
            uint32_t K, begin, ncols;
        };
        uvec<Upd> bucket(n_upd);
        parallel_ranges(n_sn, kHostGrain / 4, [&](uint32_t kb, uint32_t ke, uint32_t) {
            for (uint32_t K = kb; K < ke; ++K)
                groups_of(K, [&](uint32_t J, uint32_t t, uint32_t cnt) {
                    bucket[bump(&cursor[J + 1], shared)] = Upd{K, t, cnt};
                });
        });
        // (the cursors have advanced by one supernode: cursor[J] .. cursor[J + 1] is J's bucket now)
        P.upd_ptr.assign(cursor.begin(), cursor.begin() + n_sn + 1);
        P.upd_sn.resize(n_upd);
        P.upd_rbegin.resize(n_upd);
        P.upd_ncols.resize(n_upd);
        P.upd_rel_ptr.resize((size_t)n_upd + 1);
        P.upd_rel_ptr[0] = 0;
        parallel_ranges(n_sn, kHostGrain / 4, [&](uint32_t jb, uint32_t je, uint32_t) {
            for (uint32_t J = jb; J < je; ++J) {
                std::sort(bucket.begin() + P.upd_ptr[J], bucket.begin() + P.upd_ptr[J + 1],
                          [](const Upd& x, const Upd& y) { return x.K < y.K; });  // a K updates a J at most once
                for (uint32_t u = P.upd_ptr[J]; u < P.upd_ptr[J + 1]; ++u) {
                    P.upd_sn[u] = bucket[u].K;
                    P.upd_rbegin[u] = bucket[u].begin;
                    P.upd_ncols[u] = bucket[u].ncols;
                    P.upd_rel_ptr[u + 1] = panel_h(bucket[u].K) - bucket[u].begin;
                }
            }
        });
        {
            uint64_t total = 0;
            for (uint32_t u = 0; u < n_upd; ++u) {
                total += P.upd_rel_ptr[u + 1];
                if (total > UINT32_MAX) return;  // leave P.direct false: PCG path
                P.upd_rel_ptr[u + 1] = (uint32_t)total;
            }
        }
        P.upd_rel.resize(P.upd_rel_ptr[n_upd]);
        // where each update's block of K's panel lives: rows upd_rbegin.. to the end, all w_K columns, contiguous
        // one 32-byte record per update, what the device reads: {block start slot, T = rows in the block, w_K | ncols << 8,
        // offset of the relative positions, first column of K, 0, 0, 0}
        P.upd_rec.resize((size_t)n_upd * 8);
        std::vector<uint8_t> bad(64, 0);
        parallel_ranges(n_sn, kHostGrain / 4, [&](uint32_t jb, uint32_t je, uint32_t part) {
            for (uint32_t J = jb; J < je; ++J) {
                const uint32_t j0 = P.sn_ptr[J], wJ = panel_w(J);
                const uint32_t* jrow = P.sn_rows.data() + P.sn_row_ptr[J];
                const uint32_t hJ = panel_h(J);
                for (uint32_t u = P.upd_ptr[J]; u < P.upd_ptr[J + 1]; ++u) {
                    const uint32_t K = P.upd_sn[u], rb = P.sn_row_ptr[K], h = panel_h(K), tb = P.upd_rbegin[u], nc = P.upd_ncols[u];
                    uint32_t* rel = P.upd_rel.data() + P.upd_rel_ptr[u];
                    uint32_t pp = wJ;  // position in J's row list, below the diagonal block
                    for (uint32_t t = tb; t < h; ++t) {
                        const uint32_t row = P.sn_rows[rb + t];
                        if (t < tb + nc) rel[t - tb] = row - j0;
                        else {
                            while (pp < hJ && jrow[pp] < row) ++pp;
                            if (pp >= hJ || jrow[pp] != row) {
                                bad[part] = 1;
                                rel[t - tb] = 0;
                            } else rel[t - tb] = pp;
                        }
                    }
                    uint32_t* r = P.upd_rec.data() + (size_t)u * 8;
                    r[0] = P.panel_off[K] + tb * panel_w(K);
                    r[1] = h - tb;
                    r[2] = panel_w(K) | (nc << 8);
                    r[3] = P.upd_rel_ptr[u];
                    r[4] = P.sn_ptr[K];
                    r[5] = r[6] = r[7] = 0;
                }
            }
        });
        for (uint8_t f : bad) inconsistent = inconsistent || f;
    }
    lap("update lists");
    // ---- 6. stages: height of every supernode in the supernode tree; small panels and large panels apart ----
    std::vector<uint32_t> sn_level(n_sn, 0);
    uint32_t n_stages = 0;
    for (uint32_t s = 0; s < n_sn; ++s) {
        n_stages = std::max(n_stages, sn_level[s] + 1);
        const uint32_t last = P.sn_ptr[s + 1] - 1;
        if (parent[last] != kNone) {
            const uint32_t ps = sn_of[parent[last]];
            sn_level[ps] = std::max(sn_level[ps], sn_level[s] + 1);
        }
    }
    P.stage_ptr.assign((size_t)3 * n_stages + 1, 0);  // per stage: [thread panels | warp panels | CTA panels]
    auto bucket = [&](uint32_t s) {
        // a single thread takes panels of a few doubles that receive at most two updates; panels that do not fit a warp's
        // shared-memory stage go to whole CTAs (large.cu: TeamCaps<32>::panel); the rest to warps
        const uint32_t sz = panel_h(s) * panel_w(s);
        const bool tiny = sz <= kLanePanel && P.upd_ptr[s + 1] - P.upd_ptr[s] <= 2;
        return 3 * sn_level[s] + (tiny ? 0u : (sz <= kWarpPanelCap ? 1u : 2u));
    };
    for (uint32_t s = 0; s < n_sn; ++s) P.stage_ptr[bucket(s) + 1]++;
    for (uint32_t k = 0; k < 3 * n_stages; ++k) P.stage_ptr[k + 1] += P.stage_ptr[k];
    P.stage_sn.resize(n_sn);
    {
        std::vector<uint32_t> cur(P.stage_ptr.begin(), P.stage_ptr.end() - 1);
        for (uint32_t s = 0; s < n_sn; ++s) P.stage_sn[cur[bucket(s)]++] = s;
    }
    // Inside a stage the warp (and CTA) panels are dealt round-robin to the teams: sorted by estimated cost, descending,
    // every team receives one panel of each cost tier and the teams reach the stage's barrier together (ncu: half of the
    // warp time of lm_large_kernel was spent waiting at barriers with the panels in elimination order).
    if (const char* env = std::getenv("EZPZ_B200_STAGE_SORT"); !env || env[0] != '0') {
        std::vector<uint32_t> cost(n_sn, 0);
        for (uint32_t s = 0; s < n_sn; ++s) {
            uint64_t c = 20 + 10ull * panel_w(s) + (uint64_t)panel_h(s) * panel_w(s) / 16;
            for (uint32_t u = P.upd_ptr[s]; u < P.upd_ptr[s + 1]; ++u) {
                const uint32_t K = P.upd_sn[u];
                c += 35 + (uint64_t)(panel_h(K) - P.upd_rbegin[u]) * panel_w(K) / 16;
            }
            cost[s] = (uint32_t)std::min<uint64_t>(c, UINT32_MAX);
        }
        for (uint32_t st = 0; st < n_stages; ++st)
            for (uint32_t cls = 1; cls <= 2; ++cls)
                std::stable_sort(P.stage_sn.begin() + P.stage_ptr[3 * st + cls], P.stage_sn.begin() + P.stage_ptr[3 * st + cls + 1],
                                 [&](uint32_t x, uint32_t y) { return cost[x] > cost[y]; });
    }
    // one 32-byte record per supernode IN STAGE ORDER (what a team reads first, one memory round trip):
    // {first column, width, height, offset of its row list, panel offset, first update, number of updates, 0}
    P.stage_rec.assign((size_t)n_sn * 8, 0u);
    for (uint32_t k = 0; k < n_sn; ++k) {
        const uint32_t sn = P.stage_sn[k];
        uint32_t* r = P.stage_rec.data() + (size_t)k * 8;
        r[0] = P.sn_ptr[sn];
        r[1] = panel_w(sn);
        r[2] = panel_h(sn);
        r[3] = P.sn_row_ptr[sn];
        r[4] = P.panel_off[sn];
        r[5] = P.upd_ptr[sn];
        r[6] = P.upd_ptr[sn + 1] - P.upd_ptr[sn];
    }
    lap("stages and records");
    // ---- 7. A = JtJ: products of every entry A has, addressed to its panel slot; diagonal slots ---------------
    // Independent per column, ranges of columns on host threads: the entries A has in every column (elimination numbering)
    // with their panel slots first, then their products (fill_products).
    {
        P.diag_slot.resize(n);
        P.aent_colptr.resize((size_t)n + 1);
        P.aent_colptr[0] = 0;
        parallel_ranges(n, kHostGrain, [&](uint32_t c0, uint32_t c1, uint32_t) {
            for (uint32_t j = c0; j < c1; ++j) {
                uint32_t cnt = 0;
                for (uint32_t q = lc_ptr[j]; q < lc_ptr[j + 1]; ++q) cnt += lc_in_a[q];
                P.aent_colptr[j + 1] = cnt;
            }
        });
        prefix_sum(P.aent_colptr);
        const size_t n_ent = P.aent_colptr[n];
        P.aent_slot.resize(n_ent);
        P.aent_row.resize(n_ent);
        std::vector<uint8_t> bad(64, 0);
        parallel_ranges(n, kHostGrain, [&](uint32_t c0, uint32_t c1, uint32_t t) {
            for (uint32_t j = c0; j < c1; ++j) {
                const uint32_t J = sn_of[j], j0 = P.sn_ptr[J], j1 = P.sn_ptr[J + 1], w = j1 - j0;
                P.diag_slot[j] = P.panel_off[J] + (j - j0) * w + (j - j0);
                // rows of the panel below the diagonal block, walked together with the column's (ascending) rows
                const uint32_t* prow = P.sn_rows.data() + P.sn_row_ptr[J];
                const uint32_t ph = P.sn_row_ptr[J + 1] - P.sn_row_ptr[J];
                uint32_t pp = w, e = P.aent_colptr[j];
                for (uint32_t q = lc_ptr[j]; q < lc_ptr[j + 1]; ++q) {
                    if (!lc_in_a[q]) continue;
                    const uint32_t i = lc_row[q];
                    uint32_t pos = 0;
                    if (i < j1) pos = i - j0;
                    else {
                        while (pp < ph && prow[pp] < i) ++pp;
                        if (pp >= ph || prow[pp] != i) bad[t] = 1;
                        else pos = pp;
                    }
                    P.aent_row[e] = i;
                    P.aent_slot[e++] = P.panel_off[J] + pos * w + (j - j0);
                }
            }
        });
        for (uint8_t f : bad) inconsistent = inconsistent || f;
        fill_products(S, perm.data());
    }
    lap("products of A");
    if (inconsistent) {
        std::fprintf(stderr, "[ezpz_b200] sparse_direct: inconsistent panel structure, using the PCG path\n");
        return;
    }
    n_levels = n_stages;
    if (const char* dbg = std::getenv("EZPZ_B200_DEBUG"); dbg && dbg[0] == '1') {
        std::fprintf(stderr, "[sparse_direct] n %u nested %d supernodes %u stages %u panel doubles %u true nnz(L) %zu updates %zu\n", n,
                     (int)P.nested, n_sn, n_stages, nnz_l, nnz_l_true, P.upd_sn.size());
        for (uint32_t st = 0; st < n_stages; ++st) {
            uint32_t max_h = 0, max_u = 0, u4 = 0, u8 = 0, u16 = 0;
            uint64_t sum_u = 0;
            for (uint32_t k = P.stage_ptr[3 * st]; k < P.stage_ptr[3 * st + 3]; ++k) {
                const uint32_t sn = P.stage_sn[k], nu = P.upd_ptr[sn + 1] - P.upd_ptr[sn];
                max_h = std::max(max_h, panel_h(sn));
                max_u = std::max(max_u, nu);
                sum_u += nu;
                u4 += nu >= 4;
                u8 += nu >= 8;
                u16 += nu >= 16;
            }
            std::fprintf(stderr,
                         "  stage %u: %u thread + %u warp + %u CTA panels, tallest %u rows, most updates %u (mean %.1f; panels with >= 4 / 8 / 16 "
                         "updates: %u / %u / %u)\n",
                         st, P.stage_ptr[3 * st + 1] - P.stage_ptr[3 * st], P.stage_ptr[3 * st + 2] - P.stage_ptr[3 * st + 1],
                         P.stage_ptr[3 * st + 3] - P.stage_ptr[3 * st + 2], max_h, max_u,
                         (double)sum_u / std::max<uint32_t>(1, P.stage_ptr[3 * st + 3] - P.stage_ptr[3 * st]), u4, u8, u16);
        }
    }
    P.perm.assign(perm.begin(), perm.end());
    P.n_levels = n_levels;
    P.nnz_l = nnz_l;
    P.direct = true;
}

}  // namespace ezs
