// structure.cpp — host analysis behind ezpz_b200_structure_create (see structure.h).
//
// Steps and the reference code each one replaces:
//   1. validation              validate_variables, ezpz/src/solver.rs:142-189
//   2. rows + J pattern        Model::new, solver.rs:217-265 (pairs -> sort + dedup -> CSC); CSR = transpose
//   3. scatter slots           replaces the per-nonzero linear search of refresh_jacobian, solver.rs:412-418
//   4. pattern of A = JtJ + D  precompute_symbolic_cholesky, solver.rs:289-300 (ones-valued J, JtJ + lambda*I)
//   5. symbolic Cholesky       SymbolicLlt::try_new (faer, not in tree) -> natural-order elimination-tree
//                              column merge here (DESIGN.md §3 states the arithmetic-order spec)
//   6. connected components    of the graph of A (used to choose the large-system path)
//   7. op tapes                straight-line programme of the batched small-system kernel
#include "structure.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include "host_parallel.h"
#include "kinds.h"

using namespace ezs;

namespace {

void set_detail(ezpz_error_detail_t* d, const char* msg) {
    if (!d) return;
    std::snprintf(d->message, sizeof d->message, "%s", msg);
}

// lower(A) by columns from the patterns of J: column j holds every i >= j that shares a row with j (and j itself:
// lambda*I puts every diagonal in).  For each column its rows are walked in CSC order and their CSR entries >= j
// collected with a marker array; the short list is then sorted.
void build_a_pattern(ezpz_structure& S) {
    const uint32_t n = S.n;
    S.a_col_ptr.assign((size_t)n + 1, 0);
    S.a_row_idx.clear();
    // columns are independent: ranges of columns on host threads, each into its own list, concatenated in column order
    struct Part {
        uvec<uint32_t> rows;
        uint32_t c0 = 0, c1 = 0;
    };
    std::vector<Part> parts(16);
    uint32_t n_parts = 1;
    parallel_ranges(n, kHostGrain, [&](uint32_t c0, uint32_t c1, uint32_t t) {
        Part& part = parts[t];
        part.c0 = c0;
        part.c1 = c1;
        part.rows.reserve((size_t)(c1 - c0) * 8);
        std::vector<uint32_t> mark(n, UINT32_MAX), col;
        for (uint32_t j = c0; j < c1; ++j) {
            col.clear();
            col.push_back(j);
            mark[j] = j;
            for (uint32_t p = S.csc_col_ptr[j]; p < S.csc_col_ptr[j + 1]; ++p) {
                const uint32_t r = S.csc_row_idx[p];
                for (uint32_t q = S.csr_row_ptr[r + 1]; q-- > S.csr_row_ptr[r];) {  // columns descending: stop below j
                    const uint32_t i = S.csr_col_idx[q];
                    if (i <= j) break;
                    if (mark[i] != j) {
                        mark[i] = j;
                        col.push_back(i);
                    }
                }
            }
            std::sort(col.begin(), col.end());
            part.rows.insert(part.rows.end(), col.begin(), col.end());
            S.a_col_ptr[j + 1] = (uint32_t)col.size();  // (counts; turned into offsets below)
        }
    }, &n_parts);
    for (uint32_t j = 0; j < n; ++j) S.a_col_ptr[j + 1] += S.a_col_ptr[j];
    S.a_row_idx.resize(S.a_col_ptr[n]);
    parallel_ranges(n_parts, 1, [&](uint32_t tb, uint32_t te, uint32_t) {
        for (uint32_t t = tb; t < te; ++t)
            if (!parts[t].rows.empty()) std::copy(parts[t].rows.begin(), parts[t].rows.end(), S.a_row_idx.begin() + S.a_col_ptr[parts[t].c0]);
    });
}

// Natural-order symbolic Cholesky: struct(L_j) = struct(A_j) U (struct(L_c) \ {c}) over the children c of
// j in the elimination tree, parent(j) = min(struct(L_j) \ {j}).
void build_l_pattern(ezpz_structure& S) {
    // Column-merge symbolic Cholesky in natural order: pattern(L_j) = pattern(A_j) united with the patterns of j's children in
    // the elimination tree (minus the children themselves).  Flat arrays throughout — children as linked lists, the union
    // through a marker array, each column sorted in place: no allocation per column (two vectors per column cost ~0.2 ms of
    // malloc on a 2,000-variable system, a third of its whole analysis).
    const uint32_t n = S.n;
    constexpr uint32_t kNone = UINT32_MAX;
    std::vector<uint32_t> first_child(n, kNone), next_sibling(n, kNone), mark(n, kNone);
    S.l_col_ptr.assign((size_t)n + 1, 0);
    S.l_row_idx.clear();
    S.l_row_idx.reserve(S.a_row_idx.size() * 2);
    for (uint32_t j = 0; j < n; ++j) {
        const size_t start = S.l_row_idx.size();
        for (uint32_t p = S.a_col_ptr[j]; p < S.a_col_ptr[j + 1]; ++p) {  // sorted, the diagonal first
            mark[S.a_row_idx[p]] = j;
            S.l_row_idx.push_back(S.a_row_idx[p]);
        }
        const size_t own = S.l_row_idx.size();
        for (uint32_t c = first_child[j]; c != kNone; c = next_sibling[c])
            for (uint32_t p = S.l_col_ptr[c] + 1; p < S.l_col_ptr[c + 1]; ++p) {  // (skip c itself; the rest is >= j)
                const uint32_t row = S.l_row_idx[p];
                if (mark[row] != j) {
                    mark[row] = j;
                    S.l_row_idx.push_back(row);
                }
            }
        if (S.l_row_idx.size() > own) std::sort(S.l_row_idx.begin() + start, S.l_row_idx.end());
        S.l_col_ptr[j + 1] = (uint32_t)S.l_row_idx.size();
        if (S.l_row_idx.size() - start > 1) {  // parent = the first row below the diagonal
            const uint32_t parent = S.l_row_idx[start + 1];
            next_sibling[j] = first_child[parent];
            first_child[parent] = j;
        }
    }
}

void build_components(ezpz_structure& S) {
    const uint32_t n = S.n;
    std::vector<uint32_t> parent(n);
    std::iota(parent.begin(), parent.end(), 0u);
    auto find = [&](uint32_t v) {
        while (parent[v] != v) {
            parent[v] = parent[parent[v]];
            v = parent[v];
        }
        return v;
    };
    for (uint32_t r = 0; r < S.m; ++r) {
        const uint32_t b = S.csr_row_ptr[r], e = S.csr_row_ptr[r + 1];
        for (uint32_t p = b + 1; p < e; ++p) {
            uint32_t x = find(S.csr_col_idx[b]), y = find(S.csr_col_idx[p]);
            if (x != y) parent[std::max(x, y)] = std::min(x, y);
        }
    }
    S.comp_of.assign(n, 0);
    std::vector<uint32_t> label(n, UINT32_MAX), size;
    uint32_t nc = 0;
    for (uint32_t v = 0; v < n; ++v) {
        uint32_t root = find(v);
        if (label[root] == UINT32_MAX) {
            label[root] = nc++;
            size.push_back(0);
        }
        S.comp_of[v] = label[root];
        size[label[root]]++;
    }
    S.n_components = nc;
    S.max_component = size.empty() ? 0 : *std::max_element(size.begin(), size.end());
}

// ---- op tape -------------------------------------------------------------------------------------
struct TapeBuilder {
    std::vector<uint32_t>& t;
    uint32_t n_ops = 0;
    uint64_t n_pairs = 0;
    size_t head = 0;
    explicit TapeBuilder(std::vector<uint32_t>& tape) : t(tape) {}
    void begin(uint32_t dst, uint32_t code, uint32_t fin_kind, uint32_t fin_slot) {
        head = t.size();
        t.push_back(dst & 0xffffu);
        t.push_back((fin_slot & 0xffffu) | ((code | (fin_kind << OP_FIN_SHIFT)) << 16));
        t.push_back(0u);
        ++n_ops;
    }
    void pair(uint32_t a, uint32_t b) {
        t.push_back((a & 0xffffu) | (b << 16));
        t[head] += 1u << 16;
        ++n_pairs;
    }
    // The pairs pushed so far are the added ones (the assembly of A inside a factorisation op); lambda follows them
    // when `lambda` is set.  Pairs pushed from now on are subtracted.
    void added_done(bool lambda) {
        t[head + 2] = t[head] >> 16;
        if (lambda) t[head + 1] |= (uint32_t)OP_MID_LAMBDA << 16;
    }
};

constexpr uint32_t kMaxSymbolicVars = 200000;

// Shared-memory budget of the thread-per-problem kernel: at least one warp must fit in 227 KB.
constexpr uint32_t kMaxSmallW = (227u * 1024u) / (8u * 33u);  // 880 doubles per problem (stride 33 for 32 threads)

void build_small_program(ezpz_structure& S) {
    SmallProgram& P = S.small;
    P = SmallProgram();
    const uint32_t n = S.n, m = S.m;
    const uint32_t nnz_j = (uint32_t)S.csc_row_idx.size();
    const uint32_t nnz_l = (uint32_t)S.l_row_idx.size();
    // The A/L region doubles as the buffer of the trial point's Jacobian (the factor is dead by the time the
    // tentative step is evaluated), so it is sized for whichever is larger.
    const uint32_t lt = std::max(nnz_l, nnz_j);
    const uint64_t W = (uint64_t)n + 2ull * m + nnz_j + lt + n + S.n_side + 1;
    if (W > kMaxSmallW || n == 0) return;
    P.X0 = 0;
    P.R0 = n;
    P.RN0 = P.R0 + m;
    P.J0 = P.RN0 + m;
    P.L0 = P.J0 + nnz_j;
    P.D0 = P.L0 + lt;
    P.S0 = P.D0 + n;
    P.n_side = S.n_side;
    P.F0 = P.S0 + S.n_side;
    P.W = (uint32_t)W;
    TapeBuilder tb(P.tape);

    // position of every L entry: column-major, diagonal first; row lists for intersections
    struct RowEnt {
        uint32_t col, slot;
    };
    std::vector<std::vector<RowEnt>> lrow(n);  // strictly lower entries of row i, columns ascending
    std::vector<uint32_t> diag_slot(n);
    for (uint32_t j = 0; j < n; ++j) {
        diag_slot[j] = P.L0 + S.l_col_ptr[j];
        for (uint32_t p = S.l_col_ptr[j] + 1; p < S.l_col_ptr[j + 1]; ++p) lrow[S.l_row_idx[p]].push_back({j, P.L0 + p});
    }

    // The tape fuses what the reference does in separate passes wherever a value has exactly one consumer — same
    // operations in the same order, without the store, the reload and the second op header in between:
    //   A[i][j] = sum_r J[r][i] J[r][j] (+ lambda on the diagonal) is only ever read by the factorisation op of L[i][j]
    //   (left-looking Cholesky), so that op starts with A's products (added, from +0.0), adds lambda, and goes on
    //   subtracting the L[i][k] L[j][k];  b[i] = sum_r fma(-J[r][i], r[r]) is only read by the forward substitution op of
    //   y[i], which therefore starts with b's products.
    auto in_a = [&](uint32_t i, uint32_t j) {  // is (i, j), i >= j, an entry of lower(A)? (fill entries of L are not)
        const uint32_t* b = S.a_row_idx.data() + S.a_col_ptr[j];
        const uint32_t* e = S.a_row_idx.data() + S.a_col_ptr[j + 1];
        return std::binary_search(b, e, i);
    };
    auto a_products = [&](uint32_t i, uint32_t j, bool emit) {  // rows shared by columns i and j of J, ascending
        uint32_t count = 0;
        uint32_t pi = S.csc_col_ptr[i], pie = S.csc_col_ptr[i + 1];
        uint32_t pj = S.csc_col_ptr[j], pje = S.csc_col_ptr[j + 1];
        while (pi < pie && pj < pje) {
            const uint32_t ri = S.csc_row_idx[pi], rj = S.csc_row_idx[pj];
            if (ri == rj) {
                if (emit) tb.pair(P.J0 + pi, P.J0 + pj);
                ++count;
                ++pi;
                ++pj;
            } else if (ri < rj) ++pi;
            else ++pj;
        }
        return count;
    };
    // Opens the factorisation op of L entry (i, j) at `slot` with the assembly of A[i][j] fused in front.
    auto begin_entry = [&](uint32_t slot, uint32_t i, uint32_t j, uint32_t fin_kind, uint32_t fin_slot) {
        tb.begin(slot, 0u, fin_kind, fin_slot);
        if (i == j || in_a(i, j)) a_products(i, j, true);
        tb.added_done(i == j);
    };
    // (1)+(3) A = JtJ + lambda*I and its left-looking Cholesky, column by column: pivot, then the sub-diagonal entries
    for (uint32_t j = 0; j < n; ++j) {
        begin_entry(diag_slot[j], j, j, OP_FIN_PIVOT, 0u);
        for (const RowEnt& e : lrow[j]) tb.pair(e.slot, e.slot);
        for (uint32_t p = S.l_col_ptr[j] + 1; p < S.l_col_ptr[j + 1]; ++p) {
            const uint32_t i = S.l_row_idx[p];
            begin_entry(P.L0 + p, i, j, OP_FIN_MUL, diag_slot[j]);
            // k < j present in both row i and row j, ascending
            const auto& ri = lrow[i];
            const auto& rj = lrow[j];
            size_t a = 0, b = 0;
            while (a < ri.size() && b < rj.size() && ri[a].col < j && rj[b].col < j) {
                if (ri[a].col == rj[b].col) {
                    tb.pair(ri[a].slot, rj[b].slot);
                    ++a;
                    ++b;
                } else if (ri[a].col < rj[b].col) ++a;
                else ++b;
            }
        }
    }
    // (2)+(4) b = Jt * (-r) and the forward substitution L y = b (into d)
    for (uint32_t i = 0; i < n; ++i) {
        tb.begin(P.D0 + i, 0u, OP_FIN_MUL, diag_slot[i]);
        for (uint32_t p = S.csc_col_ptr[i]; p < S.csc_col_ptr[i + 1]; ++p) tb.pair(P.J0 + p, P.R0 + S.csc_row_idx[p]);
        for (const RowEnt& e : lrow[i]) tb.pair(e.slot, P.D0 + e.col);
    }
    // (5) backward substitution Lt d = y, rows of every column DESCENDING (on the large path this lets the
    // columns of a supernode advance together, and one order serves both paths and the oracle)
    for (uint32_t ii = n; ii-- > 0;) {
        tb.begin(P.D0 + ii, OP_INIT_DST, OP_FIN_MUL, diag_slot[ii]);
        for (uint32_t p = S.l_col_ptr[ii + 1]; p-- > S.l_col_ptr[ii] + 1;) tb.pair(P.L0 + p, P.D0 + S.l_row_idx[p]);
    }
    P.n_ops = tb.n_ops;
    P.n_pairs = tb.n_pairs;
    // a single op may not hold more than 65535 pairs (16-bit count); cannot happen below kMaxSmallW
    P.valid = true;
}


// Relative cost of evaluating one constraint of each kind (residual + Jacobian), in units of a trivial kind; only used to
// balance the constraint lists of the roles.
constexpr uint8_t kEvalCost[EZPZ_K_COUNT] = {16, 12, 8, 9, 1, 1, 1, 1, 20, 1, 1, 2, 1, 16, 16, 16, 2, 20, 12, 12, 30, 50, 30, 20, 24};
// ... in units of the tape's cost model: a flat part per constraint plus a little per unit of formula.
constexpr uint32_t eval_cost(uint32_t kind) { return 650u + 6u * kEvalCost[kind]; }

}  // namespace

namespace ezs {

// Compiles the sequential small program for R cooperating warps (see RoleBlob in structure.h).
//   * constraints: longest-processing-time partition by kEvalCost; a constraint's residual rows and Jacobian entries
//     belong to it alone, so the lists are independent.
//   * tape: ops without a dependency inside the tape (the assembly of A and of the right-hand side) are dealt over all
//     roles; the factorisation and the two substitutions go to the role that owns the connected component of their variable
//     (components of the graph of A are independent linear systems), the components being dealt over the roles by cost.
//     A single-component system therefore factorises on one role while the others wait.
//   * barriers: ops are emitted in the order of the sequential tape (a topological order); whenever an op reads or
//     overwrites a slot that ANOTHER role wrote or read since the last barrier, a barrier is put in front of it in every
//     role's tape.  Hazards inside one role are ordered by that role's program order.
void build_role_blob(const ezpz_structure& S, uint32_t R, uint32_t stride, RoleBlob& out) {
    const SmallProgram& P = S.small;
    out = RoleBlob();
    out.R = R;
    out.stride = stride;
    const uint32_t n = S.n, m = S.m, nc = S.n_cons;
    const uint32_t nnz_j = (uint32_t)S.csc_row_idx.size();
    const uint32_t sb = stride * 8u;
    // ---- constraint lists
    std::vector<std::vector<uint32_t>> clist(R);
    {
        std::vector<uint32_t> order(nc);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return eval_cost(S.cons[a].kind) > eval_cost(S.cons[b].kind); });
        std::vector<uint64_t> load(R, 0);
        for (uint32_t c : order) {
            const uint32_t r = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
            load[r] += eval_cost(S.cons[c].kind);
            clist[r].push_back(c);
        }
        for (auto& l : clist) std::sort(l.begin(), l.end());
    }
    // ---- ops of the sequential tape
    struct Op {
        uint32_t dst, fin, code, np, added, shape;
        size_t pairs;  // index of the first pair word in P.tape
        uint32_t role = 0;
        uint32_t cost() const { return 150 + 5 * np; }  // fitted, see critical_cost below
    };
    std::vector<Op> ops;
    for (size_t i = 0; i < P.tape.size();) {
        const uint32_t h0 = P.tape[i], h1 = P.tape[i + 1];
        Op o;
        o.dst = h0 & 0xffffu;
        o.np = h0 >> 16;
        o.fin = h1 & 0xffffu;
        o.code = h1 >> 16;
        o.added = P.tape[i + 2];
        o.pairs = i + 3;
        const uint32_t fk = (o.code >> OP_FIN_SHIFT) & 3u;
        o.shape = fk == OP_FIN_PIVOT ? TAPE_PIVOT : (o.code & OP_INIT_DST) ? TAPE_BACKWARD : TAPE_ENTRY;
        ops.push_back(o);
        i += 3 + o.np;
    }
    const uint32_t W = P.W;
    auto reads_of = [&](const Op& o, std::vector<uint32_t>& rd) {
        rd.clear();
        if (o.shape == TAPE_BACKWARD) rd.push_back(o.dst);
        if (o.shape != TAPE_PIVOT) rd.push_back(o.fin);
        for (uint32_t k = 0; k < o.np; ++k) {
            rd.push_back(P.tape[o.pairs + k] & 0xffffu);
            rd.push_back(P.tape[o.pairs + k] >> 16);
        }
    };
    // level 0 = no dependency on an earlier op of the tape
    std::vector<uint8_t> written(W, 0), was_read(W, 0), level0(ops.size(), 0);
    std::vector<uint32_t> rd;
    for (size_t k = 0; k < ops.size(); ++k) {
        reads_of(ops[k], rd);
        bool dep = written[ops[k].dst] || was_read[ops[k].dst];
        for (uint32_t s : rd) dep = dep || written[s];
        level0[k] = dep ? 0 : 1;
        for (uint32_t s : rd) was_read[s] = 1;
        written[ops[k].dst] = 1;
    }
    // component of the variable an op's destination belongs to
    std::vector<uint32_t> var_of_slot(W, UINT32_MAX);
    for (uint32_t j = 0; j < n; ++j) {
        var_of_slot[P.D0 + j] = j;
        for (uint32_t p = S.l_col_ptr[j]; p < S.l_col_ptr[j + 1]; ++p) var_of_slot[P.L0 + p] = j;
    }
    std::vector<uint64_t> comp_cost(std::max<uint32_t>(1, S.n_components), 0);
    for (size_t k = 0; k < ops.size(); ++k)
        if (!level0[k] && var_of_slot[ops[k].dst] != UINT32_MAX) comp_cost[S.comp_of[var_of_slot[ops[k].dst]]] += ops[k].cost();
    std::vector<uint32_t> comp_role(comp_cost.size(), 0);
    std::vector<uint64_t> load(R, 0);
    {
        std::vector<uint32_t> order(comp_cost.size());
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return comp_cost[a] > comp_cost[b]; });
        for (uint32_t c : order) {
            const uint32_t r = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
            load[r] += comp_cost[c];
            comp_role[c] = r;
        }
    }
    {   // level-0 ops on top of that, largest first, to the least loaded role
        std::vector<uint32_t> order;
        for (size_t k = 0; k < ops.size(); ++k)
            if (level0[k]) order.push_back((uint32_t)k);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ops[a].cost() > ops[b].cost(); });
        std::vector<uint64_t> l0(R, 0);
        for (uint32_t k : order) {
            const uint32_t r = (uint32_t)(std::min_element(l0.begin(), l0.end()) - l0.begin());
            l0[r] += ops[k].cost();
            ops[k].role = r;
        }
        for (uint32_t r = 0; r < R; ++r) load[r] += l0[r];
    }
    for (size_t k = 0; k < ops.size(); ++k)
        if (!level0[k]) ops[k].role = var_of_slot[ops[k].dst] == UINT32_MAX ? 0u : comp_role[S.comp_of[var_of_slot[ops[k].dst]]];
    {
        const uint64_t total = std::accumulate(load.begin(), load.end(), (uint64_t)0);
        out.busiest_share = total ? (double)*std::max_element(load.begin(), load.end()) / (double)total : 1.0;
    }
    // ---- emission with barrier insertion
    std::vector<std::vector<uint32_t>> tape(R);
    std::vector<uint32_t> n_ops(R, 0);
    constexpr uint32_t kNone = UINT32_MAX;
    std::vector<uint32_t> w_role(W, kNone), w_epoch(W, 0);  // last writer of a slot: role, epoch
    std::vector<uint32_t> r_mask(W, 0), r_epoch(W, 0);      // roles that read a slot in epoch r_epoch
    uint32_t epoch = 1;
    for (const Op& o : ops) {
        reads_of(o, rd);
        bool hazard = false;
        for (uint32_t s : rd) hazard = hazard || (w_role[s] != kNone && w_role[s] != o.role && w_epoch[s] == epoch);      // RAW
        hazard = hazard || (w_role[o.dst] != kNone && w_role[o.dst] != o.role && w_epoch[o.dst] == epoch);                // WAW
        hazard = hazard || (r_epoch[o.dst] == epoch && (r_mask[o.dst] & ~(1u << o.role)) != 0);                           // WAR
        if (hazard) {
            ++epoch;
            ++out.tape_barriers;
            for (uint32_t r = 0; r < R; ++r) {
                tape[r].insert(tape[r].end(), {0u, 0u, 0u, (uint32_t)TAPE_BARRIER | 4u << 8});
                ++n_ops[r];
            }
        }
        std::vector<uint32_t>& t = tape[o.role];
        const uint32_t n_words = 4u + 2u * o.np + ((o.np & 1u) << 1);
        t.insert(t.end(), {o.dst * sb, o.fin * sb, o.added | (o.np - o.added) << 16, o.shape | n_words << 8});
        for (uint32_t k = 0; k < o.np; ++k) {
            const uint32_t w = P.tape[o.pairs + k];
            t.push_back((w & 0xffffu) * sb);
            t.push_back((w >> 16) * sb);
        }
        if (o.np & 1u) t.insert(t.end(), {0u, 0u});  // every header on 16 bytes: the device reads headers and pairs as 128-bit words
        ++n_ops[o.role];
        for (uint32_t s : rd) {
            if (r_epoch[s] != epoch) {
                r_epoch[s] = epoch;
                r_mask[s] = 0;
            }
            r_mask[s] |= 1u << o.role;
        }
        w_role[o.dst] = o.role;
        w_epoch[o.dst] = epoch;
    }
    {   // Cost model fitted to the measured kernel times of 8 structures x 1..4 roles on B200 (4 % rms,
        // profiles/r02a_lm_small_roles.md): 150 per tape op and only ~5 per multiply-add (an op's fixed chain — header,
        // first operands, finalisation, store — dwarfs its short loop), ~650 per constraint plus ~6 per unit of formula
        // (record load, dispatch and scatter outweigh the arithmetic), 105 per element of the three folds every role
        // repeats, a few units per barrier.
        std::vector<uint64_t> tcost(R, 0), ecost(R, 0);
        for (const Op& o : ops) tcost[o.role] += o.cost();
        for (uint32_t r = 0; r < R; ++r)
            for (uint32_t c : clist[r]) ecost[r] += eval_cost(S.cons[c].kind);
        double crit = 0.0;
        for (uint32_t r = 0; r < R; ++r) crit = std::max(crit, (double)tcost[r] + (double)ecost[r]);
        out.critical_cost = crit + 105.0 * (2.0 * m + n) + (R > 1 ? 5.0 * (4 + out.tape_barriers) : 0.0);
    }
    // ---- the blob
    std::vector<uint32_t>& w = out.words;
    w.assign((size_t)kRoleHdrWords * R, 0u);
    out.cons_word = (uint32_t)w.size();
    w.resize(w.size() + (size_t)nc * (sizeof(DevCons) / 4));
    if (nc) std::memcpy(w.data() + out.cons_word, S.dev_cons.data(), (size_t)nc * sizeof(DevCons));
    auto split = [&](uint32_t total, uint32_t r, uint32_t& lo, uint32_t& hi) {
        lo = (uint32_t)((uint64_t)total * r / R);
        hi = (uint32_t)((uint64_t)total * (r + 1) / R);
    };
    for (uint32_t r = 0; r < R; ++r) {
        uint32_t* h = nullptr;
        const uint32_t cons_off = (uint32_t)w.size();
        for (uint32_t c : clist[r]) {  // constraint index + what the scatter loop needs to know about its kind
            const ezk::KindInfo& ki = ezk::kKinds[S.cons[c].kind];
            w.push_back(c | (uint32_t)ki.rows << 16 | (uint32_t)ki.emit_len[0] << 20 | (uint32_t)ki.emit_len[1] << 24);
        }
        while (w.size() & 3u) w.push_back(0u);  // tapes start on 16 bytes
        const uint32_t tape_off = (uint32_t)w.size();
        w.insert(w.end(), tape[r].begin(), tape[r].end());
        w.insert(w.end(), {0u, 0u, 0u, 0u});  // the interpreter requests the next header before it knows there is none
        h = w.data() + (size_t)kRoleHdrWords * r;
        h[0] = cons_off;
        h[1] = (uint32_t)clist[r].size();
        h[2] = tape_off;
        h[3] = n_ops[r];
        split(n, r, h[4], h[5]);
        split(m, r, h[6], h[7]);
        split(nnz_j, r, h[8], h[9]);
    }
    while (w.size() & 3u) w.push_back(0u);
}

}  // namespace ezs

namespace {

// Programme of the single-large-system path: assembly processing order, value-array layout, and the sparse
// direct schedule (sparse_direct.cpp).
void build_large_program(ezpz_structure& S, const uint32_t* order_hint = nullptr, bool hint_nested = false,
                         const ezpz_structure* same_a = nullptr) {
    LargeProgram& P = S.large;
    P = LargeProgram();
    const uint32_t n = S.n, m = S.m;
    const uint32_t nnz_j = (uint32_t)S.csc_row_idx.size();
    // Processing order of the assembly phase: tiles of kAssemblyTile consecutive input constraints, stably
    // sorted by kind inside the tile, each kind group padded to whole warps (UINT32_MAX = idle slot).  Sorting
    // the whole list by kind makes warps kind-uniform but scatters the row and slot writes (ncu on the
    // 1M-variable sketch: 595 MB of DRAM traffic for 231 MB algorithmic); plain input order keeps locality
    // but runs 6 of 32 lanes per instruction (every warp holds a dozen kinds).  Tiles give both.
    {
        // (a stable counting sort by kind per tile; tiles are independent: sizes first, then every tile fills its own span)
        constexpr uint32_t kAssemblyTile = 4096;
        const uint32_t n_tiles = (S.n_cons + kAssemblyTile - 1) / kAssemblyTile;
        uvec<uint32_t> tile_off((size_t)n_tiles + 1);
        tile_off[0] = 0;
        auto kind_counts = [&](uint32_t t, uint32_t* cnt) {
            const uint32_t t0 = t * kAssemblyTile, t1 = std::min(S.n_cons, t0 + kAssemblyTile);
            std::fill(cnt, cnt + EZPZ_K_COUNT, 0u);
            for (uint32_t c = t0; c < t1; ++c) cnt[S.cons[c].kind]++;
        };
        parallel_ranges(n_tiles, 8, [&](uint32_t tb, uint32_t te, uint32_t) {
            uint32_t cnt[EZPZ_K_COUNT];
            for (uint32_t t = tb; t < te; ++t) {
                kind_counts(t, cnt);
                uint32_t size = 0;
                for (uint32_t k = 0; k < EZPZ_K_COUNT; ++k) size += (cnt[k] + 31u) / 32u * 32u;
                tile_off[t + 1] = size;
            }
        });
        prefix_sum(tile_off);
        P.cons_order.resize(tile_off[n_tiles]);
        parallel_ranges(n_tiles, 8, [&](uint32_t tb, uint32_t te, uint32_t) {
            uint32_t cnt[EZPZ_K_COUNT], at[EZPZ_K_COUNT];
            for (uint32_t t = tb; t < te; ++t) {
                kind_counts(t, cnt);
                uint32_t off = tile_off[t];
                for (uint32_t k = 0; k < EZPZ_K_COUNT; ++k) {
                    at[k] = off;
                    off += (cnt[k] + 31u) / 32u * 32u;
                    for (uint32_t q = at[k] + cnt[k]; q < off; ++q) P.cons_order[q] = UINT32_MAX;
                }
                const uint32_t t0 = t * kAssemblyTile, t1 = std::min(S.n_cons, t0 + kAssemblyTile);
                for (uint32_t c = t0; c < t1; ++c) P.cons_order[at[S.cons[c].kind]++] = c;
            }
        });
    }
    build_sparse_direct(S, order_hint, hint_nested, same_a);
    // Direct path: J in tile order.  Every record tile owns 32 x (partials the kind emits) consecutive doubles; partial q
    // of the constraint in lane l sits at base + q * 32 + l.  A partial that accumulates into an entry the same constraint
    // already wrote (bit 31 of its slot) shares that entry's position.  Consumers that think in CSC positions (the product
    // lists of A = JtJ, the column walk of b = -Jt r, the exported Jacobian) go through jt_of_csc.
    P.jt_of_csc.clear();
    P.n_j = nnz_j;
    if (P.direct) {
        P.jt_of_csc.resize(nnz_j);
        parallel_fill(P.jt_of_csc.data(), (size_t)nnz_j, UINT32_MAX);
        const uint32_t n_rec_tiles = (uint32_t)(P.cons_order.size() / 32);
        uvec<uint64_t> tile_base((size_t)n_rec_tiles + 1);
        tile_base[0] = 0;
        for (uint32_t t = 0; t < n_rec_tiles; ++t) {  // a tile's first slot is never padding
            const ezk::KindInfo& ki = ezk::kKinds[S.cons[P.cons_order[(size_t)t * 32]].kind];
            tile_base[t + 1] = tile_base[t] + 32ull * (ki.emit_len[0] + ki.emit_len[1]);
        }
        parallel_ranges(n_rec_tiles, kHostGrain / 32, [&](uint32_t tb, uint32_t te, uint32_t) {
            for (uint32_t t = tb; t < te; ++t) {
                const ezk::KindInfo& ki = ezk::kKinds[S.cons[P.cons_order[(size_t)t * 32]].kind];
                for (uint32_t l = 0; l < 32; ++l) {
                    const uint32_t c = P.cons_order[(size_t)t * 32 + l];
                    if (c == UINT32_MAX) continue;
                    const DevCons& dc = S.dev_cons[c];
                    for (int row = 0; row < ki.rows; ++row)
                        for (int q = 0; q < ki.emit_len[row]; ++q)
                            if (!(dc.slot[row][q] & kAccumulate))
                                P.jt_of_csc[dc.slot[row][q]] =
                                    (uint32_t)(tile_base[t] + (uint64_t)((row ? ki.emit_len[0] : 0) + q) * 32 + l);
                }
            }
        });
        uint64_t base = tile_base[n_rec_tiles];
        for (uint32_t e = 0; e < nnz_j; ++e)  // pattern entries no partial ever writes stay zero: park them behind the tiles
            if (P.jt_of_csc[e] == UINT32_MAX) P.jt_of_csc[e] = (uint32_t)std::min<uint64_t>(base++, 0x7ffffffeull);
        if (base < 0x7fffffffull) {
            P.n_j = (uint32_t)base;
            parallel_ranges((uint32_t)P.aprod_a.size(), 8 * kHostGrain, [&](uint32_t b, uint32_t e, uint32_t) {
                for (uint32_t k = b; k < e; ++k) {
                    P.aprod_a[k] = P.jt_of_csc[P.aprod_a[k]];
                    P.aprod_b[k] = P.jt_of_csc[P.aprod_b[k]];
                }
            });
        } else {
            P.direct = false;  // positions must leave bit 31 free for the accumulate flag
            P.nnz_l = 0;
            P.jt_of_csc.clear();
        }
    }
    const uint64_t nnz_l = P.direct ? P.nnz_l : 0;
    const uint64_t total = (uint64_t)n + 2ull * m + P.n_j + nnz_l + 3ull * n;
    if (total >= 0xfffffff0ull) {  // 32-bit slots: fall back to the PCG path
        P.direct = false;
        P.nnz_l = 0;
        P.jt_of_csc.clear();
        P.n_j = nnz_j;
    }
    P.X0 = 0;
    P.R0 = n;
    P.RN0 = P.R0 + m;
    P.J0 = P.RN0 + m;
    P.L0 = P.J0 + P.n_j;
    P.RV0 = P.L0 + (P.direct ? P.nnz_l : 0);
    P.Y0 = P.RV0 + n;
    P.D0 = P.Y0 + n;
    P.VG = (uint64_t)P.D0 + n;
    P.built = true;
}

}  // namespace

// The analysis behind ezpz_b200_structure_create and ezpz_b200_structure_extend.  The ids of the guesses come as a list
// (`var_ids`, create) or as the membership table of an analysed structure (`present`, extend); both null = ids 0..n_vars-1.
static int32_t analyse(const ezpz_constraint_t* cons, uint32_t n_cons, const uint32_t* var_ids, const std::vector<uint8_t>* present_in,
                       uint32_t n_vars, const ezpz_structure* base, ezpz_structure_t** out, ezpz_error_detail_t* detail) {
    if (!out) return EZPZ_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (n_cons > 0 && !cons) return EZPZ_ERR_INVALID_ARGUMENT;
    // Every pass below runs over ranges of constraints (or columns) on host threads for large systems and inline on the
    // caller's thread for small ones; a pass that can fail records the FIRST offender of its range and the lowest range wins,
    // so errors are the ones a sequential walk reports.
    struct FirstBad {
        uint32_t c = UINT32_MAX, v = 0;
    };
    auto first_bad = [](const std::vector<FirstBad>& parts) {
        for (const FirstBad& f : parts)
            if (f.c != UINT32_MAX) return f;
        return FirstBad();
    };
    {
        std::vector<FirstBad> bad(16);
        parallel_ranges(n_cons, kHostGrain, [&](uint32_t cb, uint32_t ce, uint32_t t) {
            for (uint32_t c = cb; c < ce; ++c)
                if (cons[c].kind >= EZPZ_K_COUNT) {
                    bad[t].c = c;
                    return;
                }
        });
        if (const FirstBad f = first_bad(bad); f.c != UINT32_MAX) {
            set_detail(detail, "constraint kind out of range");
            if (detail) detail->constraint_id = f.c;
            return EZPZ_ERR_INVALID_ARGUMENT;
        }
    }
    // 1. validate_variables: every id named by a row list must be among the guess ids.
    std::vector<uint8_t> present;
    {
        if (var_ids) {
            uint32_t mx = 0;
            for (uint32_t k = 0; k < n_vars; ++k) mx = std::max(mx, var_ids[k]);
            present.assign((size_t)mx + 1, 0);
            for (uint32_t k = 0; k < n_vars; ++k) present[var_ids[k]] = 1;
        } else if (present_in) present = *present_in;
        const bool listed = !present.empty() || var_ids;
        std::vector<FirstBad> bad(16);
        parallel_ranges(n_cons, kHostGrain, [&](uint32_t cb, uint32_t ce, uint32_t t) {
            for (uint32_t c = cb; c < ce; ++c) {
                const ezk::KindInfo& ki = ezk::kKinds[cons[c].kind];
                for (int row = 0; row < 2; ++row)
                    for (int k = 0; k < ki.nz_len[row]; ++k) {
                        const uint32_t v = cons[c].ids[ki.nz[row][k]];
                        const bool found = listed ? (v < present.size() && present[v]) : (v < n_vars);
                        if (!found) {
                            bad[t].c = c;
                            bad[t].v = v;
                            return;
                        }
                    }
            }
        });
        if (const FirstBad f = first_bad(bad); f.c != UINT32_MAX) {
            if (detail) {
                detail->constraint_id = f.c;
                detail->variable = f.v;
                std::snprintf(detail->message, sizeof detail->message,
                              "Constraint %u references variable %u but no such variable appears in "
                              "your initial guesses.",
                              f.c, f.v);
            }
            return EZPZ_ERR_MISSING_GUESS;
        }
    }
    ezpz_structure* S = new (std::nothrow) ezpz_structure();
    if (!S) return EZPZ_ERR_INVALID_ARGUMENT;
    auto lap = [last = std::chrono::steady_clock::now()](const char* what) mutable {
        const char* dbg = std::getenv("EZPZ_B200_DEBUG");
        if (!(dbg && dbg[0] == '1')) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[structure]     %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    };
    S->n_cons = n_cons;
    S->n = n_vars;
    S->var_present.swap(present);
    S->cons.resize(n_cons);
    parallel_copy(S->cons.data(), cons, n_cons);
    // 2. rows: first row of every constraint, and the side slots (constraints with an Undefined side in input order)
    S->cons_row0.resize((size_t)n_cons + 1);
    std::vector<uint32_t> side_slot;
    S->n_side = 0;
    {
        uint32_t row_num = 0;
        for (uint32_t c = 0; c < n_cons; ++c) {
            S->cons_row0[c] = row_num;
            row_num += ezk::kKinds[cons[c].kind].rows;
            if ((cons[c].kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE || cons[c].kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE) &&
                cons[c].flags == EZPZ_SIDE_UNDEFINED) {
                if (side_slot.empty()) side_slot.assign(n_cons, UINT32_MAX);
                side_slot[c] = S->n_side++;
            }
        }
        S->cons_row0[n_cons] = row_num;
        S->m = row_num;
    }
    // Pattern of J by columns: what faer's try_new_from_indices does with a comparison sort over all named (row, variable)
    // pairs (solver.rs:255-256) is a bucket pass here — count the pairs of every column, scatter the rows into the columns'
    // buckets, sort and deduplicate each (short) bucket, compact.  Counts and cursors are bumped with relaxed atomics so that
    // ranges of constraints run on host threads; the bucket sort makes the result independent of the order of arrival.
    const uint32_t m = S->m;
    const bool shared = host_threads(n_cons, kHostGrain) > 1;
    uvec<uint32_t> col_start((size_t)n_vars + 2);
    parallel_fill(col_start.data(), col_start.size(), 0u);
    {
        std::vector<FirstBad> bad(16);
        std::vector<uint8_t> weights_one(16, 1);
        parallel_ranges(n_cons, kHostGrain, [&](uint32_t cb, uint32_t ce, uint32_t t) {
            bool one = true;
            for (uint32_t c = cb; c < ce; ++c) {
                const ezk::KindInfo& ki = ezk::kKinds[cons[c].kind];
                one = one && cons[c].weight == 1.0;
                for (int row = 0; row < ki.rows; ++row)
                    for (int k = 0; k < ki.nz_len[row]; ++k) {
                        const uint32_t v = cons[c].ids[ki.nz[row][k]];
                        if (v >= n_vars) {  // column index outside the matrix: faer's CreationError (solver.rs:256-260)
                            if (bad[t].c == UINT32_MAX) bad[t].c = c;
                            continue;
                        }
                        bump(&col_start[v + 2], shared);
                    }
            }
            weights_one[t] = one;
        });
        if (first_bad(bad).c != UINT32_MAX) {
            set_detail(detail, "Could not create matrix: index out of bounds");
            delete S;
            return EZPZ_ERR_MATRIX;
        }
        S->all_weights_one = true;
        for (uint8_t w : weights_one) S->all_weights_one = S->all_weights_one && w;
    }
    for (uint32_t j = 0; j < n_vars; ++j) col_start[j + 2] += col_start[j + 1];  // col_start[j + 1] = cursor of column j
    const size_t n_pairs = col_start[(size_t)n_vars + 1];
    uvec<uint32_t> bucket(n_pairs);
    parallel_ranges(n_cons, kHostGrain, [&](uint32_t cb, uint32_t ce, uint32_t) {
        for (uint32_t c = cb; c < ce; ++c) {
            const ezk::KindInfo& ki = ezk::kKinds[cons[c].kind];
            for (int row = 0; row < ki.rows; ++row)
                for (int k = 0; k < ki.nz_len[row]; ++k)
                    bucket[bump(&col_start[cons[c].ids[ki.nz[row][k]] + 1], shared)] = S->cons_row0[c] + row;
        }
    });
    // (the cursors have advanced by one column: col_start[j] .. col_start[j + 1] is column j's bucket now)
    S->csc_col_ptr.resize((size_t)n_vars + 1);
    S->csc_col_ptr[0] = 0;
    parallel_ranges(n_vars, kHostGrain, [&](uint32_t jb, uint32_t je, uint32_t) {
        for (uint32_t j = jb; j < je; ++j) {
            uint32_t* b = bucket.data() + col_start[j];
            uint32_t* e = bucket.data() + col_start[j + 1];
            if (!std::is_sorted(b, e)) std::sort(b, e);
            S->csc_col_ptr[j + 1] = (uint32_t)(std::unique(b, e) - b);
        }
    });
    prefix_sum(S->csc_col_ptr);
    const size_t nnz = S->csc_col_ptr[n_vars];
    S->csc_row_idx.resize(nnz);
    parallel_ranges(n_vars, kHostGrain, [&](uint32_t jb, uint32_t je, uint32_t) {
        for (uint32_t j = jb; j < je; ++j)
            std::copy(bucket.data() + col_start[j], bucket.data() + col_start[j] + (S->csc_col_ptr[j + 1] - S->csc_col_ptr[j]),
                      S->csc_row_idx.data() + S->csc_col_ptr[j]);
    });
    // CSR = transpose: a row's columns are the distinct ids of its list, ascending; its entries find their CSC positions
    // by a binary search in their (short) columns.
    S->csr_row_ptr.resize((size_t)m + 1);
    S->csr_row_ptr[0] = 0;
    auto row_cols = [&](uint32_t c, int row, uint32_t* cols) {
        const ezk::KindInfo& ki = ezk::kKinds[cons[c].kind];
        const uint32_t len = ki.nz_len[row];
        uint32_t out = 0;
        for (uint32_t k = 0; k < len; ++k) {  // insertion sort with duplicates dropped (at most 8 ids)
            const uint32_t v = cons[c].ids[ki.nz[row][k]];
            uint32_t at = out;
            while (at > 0 && cols[at - 1] > v) --at;
            if (at > 0 && cols[at - 1] == v) continue;
            for (uint32_t q = out; q > at; --q) cols[q] = cols[q - 1];
            cols[at] = v;
            ++out;
        }
        return out;
    };
    parallel_ranges(n_cons, kHostGrain, [&](uint32_t cb, uint32_t ce, uint32_t) {
        uint32_t cols[8];
        for (uint32_t c = cb; c < ce; ++c)
            for (int row = 0; row < ezk::kKinds[cons[c].kind].rows; ++row) S->csr_row_ptr[S->cons_row0[c] + row + 1] = row_cols(c, row, cols);
    });
    prefix_sum(S->csr_row_ptr);
    S->csr_col_idx.resize(nnz);
    S->csr_to_csc.resize(nnz);
    S->csc_to_csr.resize(nnz);
    parallel_ranges(n_cons, kHostGrain, [&](uint32_t cb, uint32_t ce, uint32_t) {
        uint32_t cols[8];
        for (uint32_t c = cb; c < ce; ++c)
            for (int row = 0; row < ezk::kKinds[cons[c].kind].rows; ++row) {
                const uint32_t r = S->cons_row0[c] + row, len = row_cols(c, row, cols);
                uint32_t pos = S->csr_row_ptr[r];
                for (uint32_t k = 0; k < len; ++k, ++pos) {
                    const uint32_t* b = S->csc_row_idx.data() + S->csc_col_ptr[cols[k]];
                    const uint32_t* e = S->csc_row_idx.data() + S->csc_col_ptr[cols[k] + 1];
                    const uint32_t at = (uint32_t)(std::lower_bound(b, e, r) - S->csc_row_idx.data());
                    S->csr_col_idx[pos] = cols[k];
                    S->csr_to_csc[pos] = at;
                    S->csc_to_csr[at] = pos;
                }
            }
    });
    lap("pattern of J (sort, CSC, CSR)");
    // 3. analysed constraints with scatter slots: independent per constraint
    S->dev_cons.resize(n_cons);
    parallel_ranges(n_cons, kHostGrain, [&](uint32_t cb, uint32_t ce, uint32_t) {
        for (uint32_t c = cb; c < ce; ++c) {
            const ezpz_constraint_t& src = cons[c];
            const ezk::KindInfo& ki = ezk::kKinds[src.kind];
            DevCons& dc = S->dev_cons[c];
            std::memset(&dc, 0, sizeof dc);
            dc.p0 = src.p0;
            dc.p1 = src.p1;
            dc.weight = src.weight;
            dc.kind = src.kind;
            dc.flags = src.flags;
            dc.row0 = S->cons_row0[c];
            dc.side_slot = side_slot.empty() ? UINT32_MAX : side_slot[c];
            std::memcpy(dc.ids, src.ids, sizeof dc.ids);
            for (int row = 0; row < ki.rows; ++row) {
                const uint32_t r = dc.row0 + row;
                for (int k = 0; k < ki.emit_len[row]; ++k) {
                    const uint32_t col = src.ids[ki.emit[row][k]];
                    const uint32_t* b = S->csc_row_idx.data() + S->csc_col_ptr[col];
                    const uint32_t* e = S->csc_row_idx.data() + S->csc_col_ptr[col + 1];
                    const uint32_t* it = std::lower_bound(b, e, r);
                    uint32_t slot = (uint32_t)(it - S->csc_row_idx.data());
                    for (int q = 0; q < k; ++q)
                        if ((dc.slot[row][q] & ~kAccumulate) == slot) {
                            slot |= kAccumulate;
                            break;
                        }
                    dc.slot[row][k] = slot;
                }
            }
        }
    });
    lap("scatter slots");
    build_a_pattern(*S);
    lap("pattern of A");
    // The natural-order symbolic factorisation feeds the tape of the batched kernel.  A system whose state cannot fit that
    // kernel whatever the fill (W >= 3n + 2m + nnz(J) + nnz(A), build_small_program) takes the large path, which orders
    // and factorises on its own (sparse_direct.cpp): its natural L pattern is computed only if somebody asks
    // (ezpz_b200_structure_dims / _pattern_a; on a 20,000-variable lattice it cost a third of the analysis).
    S->have_l_pattern = n_vars <= kMaxSymbolicVars;
    const bool may_be_small = 3ull * n_vars + 2ull * S->m + nnz + std::max(S->a_row_idx.size(), nnz) + S->n_side + 1 <= kMaxSmallW;
    if (S->have_l_pattern && may_be_small) {
        build_l_pattern(*S);
        S->l_pattern_built = true;
    }
    build_components(*S);
    lap("natural L pattern, components");
    if (S->l_pattern_built) build_small_program(*S);
    if (!S->small.valid) {
        // constraints added to an analysed structure of the sparse-direct path (ezpz_b200_structure_extend): its elimination
        // order is kept, and its whole schedule when A did not gain an entry (EZPZ_B200_EXTEND_FULL=1: always re-derive it)
        const bool keep = base && base->large.built && base->large.direct && base->large.perm.size() == n_vars;
        const char* full = std::getenv("EZPZ_B200_EXTEND_FULL");
        const bool same_a = keep && !(full && full[0] == '1') && base->a_col_ptr == S->a_col_ptr && base->a_row_idx == S->a_row_idx;
        build_large_program(*S, keep ? base->large.perm.data() : nullptr, keep && base->large.nested, same_a ? base : nullptr);
    }
    lap("programme");
    *out = S;
    return EZPZ_OK;
}

extern "C" {

int32_t ezpz_b200_structure_create(const ezpz_constraint_t* cons, uint32_t n_cons, const uint32_t* var_ids,
                                   uint32_t n_vars, ezpz_structure_t** out, ezpz_error_detail_t* detail) {
    return analyse(cons, n_cons, var_ids, nullptr, n_vars, nullptr, out, detail);
}

int32_t ezpz_b200_structure_extend(const ezpz_structure_t* base, const ezpz_constraint_t* extra, uint32_t n_extra,
                                   ezpz_structure_t** out, ezpz_error_detail_t* detail) {
    if (!out) return EZPZ_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!base || (n_extra > 0 && !extra) || (uint64_t)base->n_cons + n_extra > UINT32_MAX) return EZPZ_ERR_INVALID_ARGUMENT;
    uvec<ezpz_constraint_t> all((size_t)base->n_cons + n_extra);
    parallel_copy(all.data(), base->cons.data(), base->n_cons);
    std::copy(extra, extra + n_extra, all.data() + base->n_cons);
    return analyse(all.data(), (uint32_t)all.size(), nullptr, &base->var_present, base->n, base, out, detail);
}

void ezpz_b200_structure_destroy(ezpz_structure_t* s) {
    if (!s) return;
    release_device_copies(s);
    for (ezs::RoleBlob* p : s->role_probes) delete p;
    delete s;
}

// The natural-order L pattern of a structure that did not need it at creation, on first request.
static void ensure_l_pattern(const ezpz_structure_t* cs) {
    ezpz_structure* s = const_cast<ezpz_structure*>(cs);
    std::lock_guard<std::mutex> lock(s->dev_mutex);
    if (s->l_pattern_built) return;
    if (s->have_l_pattern) build_l_pattern(*s);
    else {
        s->l_col_ptr.assign((size_t)s->n + 1, 0);
        s->l_row_idx.clear();
    }
    s->l_pattern_built = true;
}

int32_t ezpz_b200_structure_dims(const ezpz_structure_t* s, uint32_t* m, uint32_t* n, uint64_t* nnz_j,
                                 uint64_t* nnz_a, uint64_t* nnz_l, uint32_t* n_components) {
    if (!s) return EZPZ_ERR_INVALID_ARGUMENT;
    if (nnz_l) ensure_l_pattern(s);
    if (m) *m = s->m;
    if (n) *n = s->n;
    if (nnz_j) *nnz_j = s->csc_row_idx.size();
    if (nnz_a) *nnz_a = s->a_row_idx.size();
    if (nnz_l) *nnz_l = s->l_row_idx.size();
    if (n_components) *n_components = s->n_components;
    return EZPZ_OK;
}

int32_t ezpz_b200_structure_pattern(const ezpz_structure_t* s, const uint32_t** csc_col_ptr,
                                    const uint32_t** csc_row_idx, const uint32_t** csr_row_ptr,
                                    const uint32_t** csr_col_idx) {
    if (!s) return EZPZ_ERR_INVALID_ARGUMENT;
    if (csc_col_ptr) *csc_col_ptr = s->csc_col_ptr.data();
    if (csc_row_idx) *csc_row_idx = s->csc_row_idx.data();
    if (csr_row_ptr) *csr_row_ptr = s->csr_row_ptr.data();
    if (csr_col_idx) *csr_col_idx = s->csr_col_idx.data();
    return EZPZ_OK;
}

int32_t ezpz_b200_structure_pattern_a(const ezpz_structure_t* s, const uint32_t** a_col_ptr,
                                      const uint32_t** a_row_idx, const uint32_t** l_col_ptr,
                                      const uint32_t** l_row_idx) {
    if (!s) return EZPZ_ERR_INVALID_ARGUMENT;
    if (l_col_ptr || l_row_idx) ensure_l_pattern(s);
    if (a_col_ptr) *a_col_ptr = s->a_col_ptr.data();
    if (a_row_idx) *a_row_idx = s->a_row_idx.data();
    if (l_col_ptr) *l_col_ptr = s->l_col_ptr.data();
    if (l_row_idx) *l_row_idx = s->l_row_idx.data();
    return EZPZ_OK;
}

int32_t ezpz_b200_structure_ordering(const ezpz_structure_t* s, int32_t* path, const uint32_t** elim_order,
                                     int32_t* nested, uint32_t* n_levels, uint64_t* nnz_l, uint32_t* sum_chunk) {
    if (!s) return EZPZ_ERR_INVALID_ARGUMENT;
    const LargeProgram& P = s->large;
    const bool direct = P.built && P.direct;
    if (path) *path = !P.built ? 0 : (P.direct ? 1 : 2);
    if (elim_order) *elim_order = direct ? P.perm.data() : nullptr;
    if (nested) *nested = direct && P.nested ? 1 : 0;
    if (n_levels) *n_levels = direct ? P.n_levels : 0;
    if (nnz_l) *nnz_l = direct ? P.nnz_l : 0;
    // large.cu: single-CTA systems (n + m + nnz <= 4,096) fold sequentially, larger ones (cluster or grid) in chunks of 1,024
    if (sum_chunk) *sum_chunk = P.built ? ezs::sum_chunk_for(s->n, s->m, s->csc_row_idx.size()) : 0u;
    return EZPZ_OK;
}

int32_t ezpz_b200_structure_role_program(const ezpz_structure_t* s, uint32_t roles, uint32_t stride, uint32_t* words,
                                         uint64_t cap, uint64_t* n_words, uint32_t* cons_word, uint32_t* dims) {
    if (!s || roles == 0 || roles > 8 || stride == 0) return EZPZ_ERR_INVALID_ARGUMENT;
    if (!s->small.valid) return EZPZ_ERR_UNSUPPORTED;
    RoleBlob blob;
    build_role_blob(*s, roles, stride, blob);
    if (n_words) *n_words = blob.words.size();
    if (cons_word) *cons_word = blob.cons_word;
    if (dims) {
        dims[0] = s->small.W;
        dims[1] = s->n_cons;
        dims[2] = blob.tape_barriers;
        dims[3] = (uint32_t)blob.critical_cost;
    }
    if (words) std::memcpy(words, blob.words.data(), sizeof(uint32_t) * (size_t)std::min<uint64_t>(cap, blob.words.size()));
    return EZPZ_OK;
}

int32_t ezpz_b200_structure_rows(const ezpz_structure_t* s, const uint32_t** cons_row0) {
    if (!s || !cons_row0) return EZPZ_ERR_INVALID_ARGUMENT;
    *cons_row0 = s->cons_row0.data();
    return EZPZ_OK;
}

// Hash of everything the host analysis produced (patterns, scatter slots, tapes, the large programme): two structures with
// the same fingerprint drive the device through the same arithmetic.  Used by the tests of the threaded analysis phases and
// of ezpz_b200_structure_extend.
uint64_t ezpz_b200_structure_fingerprint(const ezpz_structure_t* s) {
    if (!s) return 0;
    uint64_t h = 0x9e3779b97f4a7c15ull;
    auto mix = [&](uint64_t v) {
        h ^= v;
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 29;
    };
    auto vec = [&](const auto& v) {
        mix(v.size());
        for (uint32_t x : v) mix(x);
    };
    mix(s->n_cons), mix(s->n), mix(s->m), mix(s->n_side), mix(s->all_weights_one), mix(s->have_l_pattern);
    vec(s->cons_row0), vec(s->csc_col_ptr), vec(s->csc_row_idx), vec(s->csr_row_ptr), vec(s->csr_col_idx);
    vec(s->csr_to_csc), vec(s->csc_to_csr), vec(s->a_col_ptr), vec(s->a_row_idx);
    if (s->l_pattern_built) vec(s->l_col_ptr), vec(s->l_row_idx);  // (computed on request for systems of the large path)
    vec(s->comp_of), mix(s->n_components), mix(s->max_component);
    mix(s->dev_cons.size());
    for (const DevCons& dc : s->dev_cons) {
        uint64_t w[sizeof(DevCons) / 8];
        std::memcpy(w, &dc, sizeof dc);
        for (uint64_t x : w) mix(x);
    }
    const SmallProgram& sp = s->small;
    mix(sp.valid), mix(sp.W), mix(sp.X0), mix(sp.R0), mix(sp.RN0), mix(sp.J0), mix(sp.L0), mix(sp.D0), mix(sp.S0), mix(sp.F0);
    mix(sp.n_side), mix(sp.n_ops), mix(sp.n_pairs), vec(sp.tape);
    const LargeProgram& P = s->large;
    mix(P.built), mix(P.direct), mix(P.nested), mix(P.X0), mix(P.R0), mix(P.RN0), mix(P.J0), mix(P.L0), mix(P.RV0), mix(P.Y0);
    mix(P.D0), mix(P.VG), mix(P.n_levels), mix(P.nnz_l), mix(P.n_j);
    vec(P.cons_order), vec(P.perm), vec(P.jt_of_csc), vec(P.sn_ptr), vec(P.sn_row_ptr), vec(P.sn_rows), vec(P.panel_off);
    vec(P.upd_ptr), vec(P.upd_sn), vec(P.upd_rbegin), vec(P.upd_ncols), vec(P.upd_rel_ptr), vec(P.upd_rel), vec(P.upd_rec);
    vec(P.stage_ptr), vec(P.stage_sn), vec(P.stage_rec), vec(P.aent_slot), vec(P.aprod_ptr), vec(P.aprod_a), vec(P.aprod_b);
    vec(P.diag_slot), vec(P.aent_colptr), vec(P.aent_row);
    return h;
}

void ezpz_b200_shard_range(uint64_t batch, uint32_t rank, uint32_t world, uint64_t* begin, uint64_t* end) {
    if (world == 0) world = 1;
    if (rank >= world) rank = world - 1;
    const uint64_t base = batch / world, extra = batch % world;
    const uint64_t b = (uint64_t)rank * base + std::min<uint64_t>(rank, extra);
    const uint64_t len = base + (rank < extra ? 1 : 0);
    if (begin) *begin = b;
    if (end) *end = b + len;
}

void ezpz_b200_config_default(ezpz_config_t* cfg) {
    if (!cfg) return;
    cfg->max_iterations = 35;       // solver.rs:73-80
    cfg->residual_tolerance = 1e-8;
    cfg->step_tolerance = 1e-12;
    cfg->initial_lambda = 1e-9;
}

uint32_t ezpz_b200_abi_version(void) { return EZPZ_B200_ABI_VERSION; }

const char* ezpz_b200_status_name(int32_t status) {
    switch (status) {
        case EZPZ_OK: return "OK";
        case EZPZ_ERR_NOT_FOUND: return "NotFound";
        case EZPZ_ERR_WRONG_NUMBER_GUESSES: return "WrongNumberGuesses";
        case EZPZ_ERR_MISSING_GUESS: return "MissingGuess";
        case EZPZ_ERR_MATRIX: return "FaerMatrix";
        case EZPZ_ERR_FAER: return "Faer";
        case EZPZ_ERR_SOLVE: return "FaerSolve";
        case EZPZ_ERR_SVD: return "FaerSvd";
        case EZPZ_ERR_EMPTY_SYSTEM: return "EmptySystemNotAllowed";
        case EZPZ_ERR_INVALID_ARGUMENT: return "InvalidArgument";
        case EZPZ_ERR_NO_DEVICE: return "NoDevice";
        case EZPZ_ERR_CUDA: return "Cuda";
        case EZPZ_ERR_UNSUPPORTED: return "Unsupported";
        case EZPZ_ERR_TOO_LARGE: return "TooLarge";
        case EZPZ_ERR_PARSE: return "Parse";
        case EZPZ_ERR_TEXT_MISSING_GUESS: return "TextMissingGuess";
        case EZPZ_ERR_TEXT_UNUSED_GUESSES: return "TextUnusedGuesses";
        case EZPZ_ERR_TEXT_UNDEFINED_POINT: return "TextUndefinedPoint";
        default: return "Unknown";
    }
}

}  // extern "C"
