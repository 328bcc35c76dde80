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
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <vector>

#include "structure.h"

namespace ezs {

namespace {

constexpr uint32_t kNone = UINT32_MAX;
constexpr uint32_t kNaturalMaxHeight = 192;     // natural order is kept when its elimination tree is this shallow
constexpr uint32_t kLeafVars = 48;              // dissection stops at parts of this many variables
constexpr uint64_t kMaxFactorEntries = 1ull << 29;  // beyond this nnz(L) the PCG path is used
constexpr uint32_t kSoloEntries = 32;           // a level with at most this many work items is run by one CTA (its
                                                // 16 warps take one or two items each), sparing a grid barrier

struct Graph {
    std::vector<uint32_t> ptr, adj;  // symmetric adjacency without self loops
};

Graph build_graph(const ezpz_structure& S) {
    const uint32_t n = S.n;
    Graph g;
    g.ptr.assign((size_t)n + 1, 0);
    for (uint32_t j = 0; j < n; ++j)
        for (uint32_t p = S.a_col_ptr[j]; p < S.a_col_ptr[j + 1]; ++p) {
            const uint32_t i = S.a_row_idx[p];
            if (i == j) continue;
            g.ptr[i + 1]++;
            g.ptr[j + 1]++;
        }
    for (uint32_t v = 0; v < n; ++v) g.ptr[v + 1] += g.ptr[v];
    g.adj.resize(g.ptr[n]);
    std::vector<uint32_t> cur(g.ptr.begin(), g.ptr.end() - 1);
    for (uint32_t j = 0; j < n; ++j)
        for (uint32_t p = S.a_col_ptr[j]; p < S.a_col_ptr[j + 1]; ++p) {
            const uint32_t i = S.a_row_idx[p];
            if (i == j) continue;
            g.adj[cur[i]++] = j;
            g.adj[cur[j]++] = i;
        }
    // ascending neighbour lists (columns are visited in ascending j, rows ascending inside a column)
    for (uint32_t v = 0; v < n; ++v) std::sort(g.adj.begin() + g.ptr[v], g.adj.begin() + g.ptr[v + 1]);
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

}  // namespace

// Fills the sparse-direct part of S.large.  Leaves P.direct false when the factor would be too large.
void build_sparse_direct(ezpz_structure& S) {
    LargeProgram& P = S.large;
    const uint32_t n = S.n;
    P.direct = false;
    if (n == 0) return;
    const char* force = std::getenv("EZPZ_B200_FORCE_PCG");
    if (force && force[0] == '1') return;
    const Graph g = build_graph(S);

    // ---- 2. elimination order ---------------------------------------------------------------------
    std::vector<uint32_t> perm(n), iperm(n), parent, level;
    std::iota(perm.begin(), perm.end(), 0u);
    std::iota(iperm.begin(), iperm.end(), 0u);
    uint32_t n_levels = etree_levels(g, perm, iperm, parent, level);
    P.nested = false;
    const char* force_nd = std::getenv("EZPZ_B200_FORCE_ND");
    if (n_levels > kNaturalMaxHeight || (force_nd && force_nd[0] == '1')) {
        std::vector<uint32_t> nd, ind(n), nparent, nlevel;
        nested_dissection(g, n, nd);
        for (uint32_t j = 0; j < n; ++j) ind[nd[j]] = j;
        const uint32_t h = etree_levels(g, nd, ind, nparent, nlevel);
        if (h < n_levels || (force_nd && force_nd[0] == '1')) {
            perm.swap(nd);
            iperm.swap(ind);
            parent.swap(nparent);
            level.swap(nlevel);
            n_levels = h;
            P.nested = true;
        }
    }

    // ---- 3. symbolic factorisation, by columns -----------------------------------------------------
    // struct(L_j) = {i > j : A_perm(i, j) != 0}  U  (struct(L_c) \ {j}) over the etree children c of j.
    std::vector<uint32_t> lc_ptr((size_t)n + 1, 0), lc_row;
    std::vector<uint8_t> lc_in_a;
    lc_row.reserve((size_t)g.adj.size() * 2);
    lc_in_a.reserve((size_t)g.adj.size() * 2);
    {
        std::vector<uint32_t> child_head(n, kNone), child_next(n, kNone), mark(n, kNone);
        std::vector<std::pair<uint32_t, uint8_t>> col;
        for (uint32_t j = 0; j < n; ++j) {
            col.clear();
            mark[j] = j;
            const uint32_t v = perm[j];
            for (uint32_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                const uint32_t i = iperm[g.adj[p]];
                if (i > j && mark[i] != j) {
                    mark[i] = j;
                    col.push_back({i, 1});
                }
            }
            for (uint32_t c = child_head[j]; c != kNone; c = child_next[c])
                for (uint32_t p = lc_ptr[c]; p < lc_ptr[c + 1]; ++p) {
                    const uint32_t i = lc_row[p];
                    if (mark[i] != j) {  // i == j is marked already
                        mark[i] = j;
                        col.push_back({i, 0});
                    }
                }
            std::sort(col.begin(), col.end());
            for (const auto& e : col) {
                lc_row.push_back(e.first);
                lc_in_a.push_back(e.second);
            }
            lc_ptr[j + 1] = (uint32_t)lc_row.size();
            if (lc_row.size() > kMaxFactorEntries) return;
            if (!col.empty()) {  // parent(j) = first sub-diagonal row
                const uint32_t par = col.front().first;
                child_next[j] = child_head[par];
                child_head[par] = j;
            }
        }
    }
    const uint32_t nnz_l = (uint32_t)lc_row.size();

    // ---- 4. L by rows (value order), level-ordered column work lists --------------------------------
    P.lr_ptr.assign((size_t)n + 1, 0);
    for (uint32_t q = 0; q < nnz_l; ++q) P.lr_ptr[lc_row[q] + 1]++;
    for (uint32_t i = 0; i < n; ++i) P.lr_ptr[i + 1] += P.lr_ptr[i];
    P.lr_col.resize(nnz_l);
    std::vector<uint32_t> slot_of(nnz_l);  // CSC entry -> position in row order
    {
        std::vector<uint32_t> cur(P.lr_ptr.begin(), P.lr_ptr.end() - 1);
        for (uint32_t j = 0; j < n; ++j)  // columns ascending => every row fills with ascending columns
            for (uint32_t q = lc_ptr[j]; q < lc_ptr[j + 1]; ++q) {
                const uint32_t pos = cur[lc_row[q]]++;
                P.lr_col[pos] = j;
                slot_of[q] = pos;
            }
    }
    P.lvl_ptr.assign((size_t)n_levels + 1, 0);
    for (uint32_t j = 0; j < n; ++j) P.lvl_ptr[level[j] + 1]++;
    for (uint32_t l = 0; l < n_levels; ++l) P.lvl_ptr[l + 1] += P.lvl_ptr[l];
    P.lvl_cols.resize(n);
    {
        std::vector<uint32_t> cur(P.lvl_ptr.begin(), P.lvl_ptr.end() - 1);
        for (uint32_t j = 0; j < n; ++j) P.lvl_cols[cur[level[j]]++] = j;
    }
    P.ent_ptr.assign((size_t)n + 1, 0);
    P.ent_row.resize(nnz_l);
    P.ent_col.resize(nnz_l);
    P.ent_slot.resize(nnz_l);
    {
        uint32_t run = 0;
        for (uint32_t p = 0; p < n; ++p) {
            const uint32_t j = P.lvl_cols[p];
            P.ent_ptr[p] = run;
            for (uint32_t q = lc_ptr[j]; q < lc_ptr[j + 1]; ++q, ++run) {
                P.ent_row[run] = lc_row[q];
                P.ent_col[run] = j;
                P.ent_slot[run] = slot_of[q] | (lc_in_a[q] ? kEntryInA : 0u);
            }
        }
        P.ent_ptr[n] = run;
    }
    P.lvl_maxrow.assign(n_levels, 0);
    for (uint32_t l = 0; l < n_levels; ++l)
        for (uint32_t p = P.lvl_ptr[l]; p < P.lvl_ptr[l + 1]; ++p)
            P.lvl_maxrow[l] = std::max(P.lvl_maxrow[l], P.lr_ptr[P.lvl_cols[p] + 1] - P.lr_ptr[P.lvl_cols[p]]);
    // first level from which every remaining level is small enough for one CTA
    P.solo_level = n_levels;
    while (P.solo_level > 0) {
        const uint32_t l = P.solo_level - 1;
        const uint64_t items = (uint64_t)(P.lvl_ptr[l + 1] - P.lvl_ptr[l]) + (P.ent_ptr[P.lvl_ptr[l + 1]] - P.ent_ptr[P.lvl_ptr[l]]);
        if (items > kSoloEntries) break;
        --P.solo_level;
    }
    // static row intersections (see structure.h), built in parallel over entries
    {
        P.ent_mask_ptr.assign((size_t)nnz_l + 1, 0);
        uint64_t words = 0;
        for (uint32_t e = 0; e < nnz_l; ++e) {
            const uint32_t i = P.ent_row[e], j = P.ent_col[e], s = P.ent_slot[e] & ~kEntryInA;
            P.ent_mask_ptr[e] = (uint32_t)words;
            words += (P.lr_ptr[j + 1] - P.lr_ptr[j] + 31) / 32 + (s - P.lr_ptr[i] + 31) / 32;
            if (words >= 0xffffffffull) return;  // 32-bit offsets: leave P.direct false (PCG path)
        }
        P.ent_mask_ptr[nnz_l] = (uint32_t)words;
        P.ent_mask.assign(words, 0u);
        auto fill = [&](uint32_t e0, uint32_t e1) {
            for (uint32_t e = e0; e < e1; ++e) {
                const uint32_t i = P.ent_row[e], j = P.ent_col[e], s = P.ent_slot[e] & ~kEntryInA;
                const uint32_t rj0 = P.lr_ptr[j], len_j = P.lr_ptr[j + 1] - rj0, ri0 = P.lr_ptr[i], pre_i = s - ri0;
                uint32_t* mj = P.ent_mask.data() + P.ent_mask_ptr[e];
                uint32_t* mi = mj + (len_j + 31) / 32;
                uint32_t a = 0, b = 0;
                while (a < pre_i && b < len_j) {
                    const uint32_t ca = P.lr_col[ri0 + a], cb = P.lr_col[rj0 + b];
                    if (ca == cb) {
                        mi[a >> 5] |= 1u << (a & 31);
                        mj[b >> 5] |= 1u << (b & 31);
                        ++a;
                        ++b;
                    } else if (ca < cb) ++a;
                    else ++b;
                }
            }
        };
        const uint32_t nt = nnz_l < (1u << 16) ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (nt == 1) fill(0, nnz_l);
        else {
            std::vector<std::thread> pool;
            for (uint32_t t = 0; t < nt; ++t)
                pool.emplace_back(fill, (uint32_t)((uint64_t)nnz_l * t / nt), (uint32_t)((uint64_t)nnz_l * (t + 1) / nt));
            for (auto& th : pool) th.join();
        }
    }
    // products of A = JtJ for the entries A itself has
    {
        P.aent.clear();
        P.aprod_ptr.assign(1, 0u);
        P.aprod_a.clear();
        P.aprod_b.clear();
        for (uint32_t e = 0; e < nnz_l; ++e) {
            if (!(P.ent_slot[e] & kEntryInA)) continue;
            const uint32_t ci = perm[P.ent_row[e]], cj = perm[P.ent_col[e]];
            uint32_t pi = S.csc_col_ptr[ci], pj = S.csc_col_ptr[cj];
            const uint32_t pie = S.csc_col_ptr[ci + 1], pje = S.csc_col_ptr[cj + 1];
            while (pi < pie && pj < pje) {
                const uint32_t ri = S.csc_row_idx[pi], rj = S.csc_row_idx[pj];
                if (ri == rj) {
                    P.aprod_a.push_back(pi);
                    P.aprod_b.push_back(pj);
                    ++pi;
                    ++pj;
                } else if (ri < rj) ++pi;
                else ++pj;
            }
            P.aent.push_back(e);
            P.aprod_ptr.push_back((uint32_t)P.aprod_a.size());
        }
    }
    if (const char* dbg = std::getenv("EZPZ_B200_DEBUG"); dbg && dbg[0] == '1') {
        std::fprintf(stderr, "[sparse_direct] n %u nnz_l %u levels %u solo_level %u nested %d\n", n, nnz_l, n_levels,
                     P.solo_level, (int)P.nested);
        for (uint32_t l = 0; l < n_levels; ++l) {
            uint32_t maxrow = 0;
            for (uint32_t p = P.lvl_ptr[l]; p < P.lvl_ptr[l + 1]; ++p)
                maxrow = std::max(maxrow, P.lr_ptr[P.lvl_cols[p] + 1] - P.lr_ptr[P.lvl_cols[p]]);
            std::fprintf(stderr, "  level %u: cols %u entries %u max row length %u\n", l, P.lvl_ptr[l + 1] - P.lvl_ptr[l],
                         P.ent_ptr[P.lvl_ptr[l + 1]] - P.ent_ptr[P.lvl_ptr[l]], maxrow);
        }
    }
    P.perm = perm;
    P.n_levels = n_levels;
    P.nnz_l = nnz_l;
    P.direct = true;
}

}  // namespace ezs
