// host_api.cpp — ezpz::solve / ezpz::solve_analysis behind the C ABI (ezpz_b200_solve).
//
// Host control that surrounds the hot path in the reference and stays host-side here:
//   * the empty-request short circuit               ezpz/src/lib.rs:155-170
//   * the priority loop over ascending levels       lib.rs:199-246 (always restarting from the ORIGINAL guesses)
//   * the static lint                               ezpz/src/warnings.rs:34-59
//   * result packing (unsatisfied in original request indices, priority_solved, warnings)  lib.rs:333-355
// Everything numeric — side inference, the LM loop, the unsatisfied check, the freedom analysis — runs on
// the device through ezpz_b200_structure_create + ezpz_b200_solve_one + ezpz_b200_freedom_analysis.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/ezpz_b200.h"
#include "device.h"
#include "kinds.h"

namespace {

// ---- topology cache -------------------------------------------------------------------------------------------
// The reference analyses the system inside every solve call (Model::new, lib.rs:279).  Interactive callers solve the
// same sketch topology again and again (dragging a point changes guesses, not constraints), so ezpz_b200_solve keeps
// the analysed structures of the last few distinct constraint lists per context: a repeated call skips the pattern
// build, the symbolic factorisation, the tape / supernode schedule and the upload of the device tables.
struct CachedStructure {
    uint64_t hash = 0;
    std::vector<ezpz_constraint_t> cons;
    std::vector<uint32_t> var_ids;
    bool has_var_ids = false;
    uint32_t n_vars = 0;
    ezpz_structure_t* S = nullptr;
    uint64_t last_use = 0;
};
struct StructureCache {
    std::vector<CachedStructure> entries;
    uint64_t clock = 0, hits = 0, misses = 0, extended = 0;
};
constexpr size_t kCacheEntries = 8;
constexpr uint32_t kCacheMaxVars = 1u << 18;  // larger systems hold GBs of device tables: not cached
// An analysed structure keeps device tables in proportion to its size (hundreds of bytes per variable on the large path):
// the cache is bounded by the variables it holds, not only by its entry count.
constexpr uint64_t kCacheVarBudget = 1u << 19;

// Multiply-xorshift over 64-bit words (the records are 64 bytes each, the id lists are padded by the tail loop): one
// multiply per 8 bytes instead of FNV-1a's one per byte, which cost ~130 us on the 128 KB of massive_parallel_system.
uint64_t mix_hash(const void* data, size_t bytes, uint64_t h) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 32;
    }
    uint64_t tail = 0;
    if (i < bytes) {
        std::memcpy(&tail, p + i, bytes - i);
        h = (h ^ tail) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 32;
    }
    return h;
}

// Returns a structure for this constraint list (created on a miss); *owned = true when the caller must destroy it.
int32_t cached_structure(ezpz_context_t* ctx, const std::vector<ezpz_constraint_t>& cons, const uint32_t* var_ids, uint32_t n_vars,
                         ezpz_structure_t** out, bool* owned, ezpz_error_detail_t* detail) {
    *owned = true;
    const bool disabled = [] {
        const char* e = std::getenv("EZPZ_B200_NO_STRUCTURE_CACHE");
        return e && e[0] == '1';
    }();
    if (!ctx || n_vars > kCacheMaxVars || disabled)
        return ezpz_b200_structure_create(cons.data(), (uint32_t)cons.size(), var_ids, n_vars, out, detail);
    if (!ctx->structure_cache) ctx->structure_cache = new StructureCache();
    StructureCache& C = *static_cast<StructureCache*>(ctx->structure_cache);
    uint64_t h = mix_hash(cons.data(), cons.size() * sizeof(ezpz_constraint_t), 0xcbf29ce484222325ull);
    h = mix_hash(&n_vars, sizeof n_vars, h);
    if (var_ids) h = mix_hash(var_ids, n_vars * sizeof(uint32_t), h);
    for (CachedStructure& e : C.entries) {
        if (e.hash != h || e.n_vars != n_vars || e.cons.size() != cons.size() || e.has_var_ids != (var_ids != nullptr)) continue;
        if (std::memcmp(e.cons.data(), cons.data(), cons.size() * sizeof(ezpz_constraint_t)) != 0) continue;
        if (var_ids && std::memcmp(e.var_ids.data(), var_ids, n_vars * sizeof(uint32_t)) != 0) continue;
        e.last_use = ++C.clock;
        ++C.hits;
        *out = e.S;
        *owned = false;
        return EZPZ_OK;
    }
    ++C.misses;
    // A miss whose list continues a cached one (a constraint added to a sketch that was just solved: the trim and drag
    // workflows, tests.rs:748-897) extends that structure instead of analysing from scratch: same patterns and tapes, and a
    // sparse-direct system keeps its elimination order (ezpz_b200_structure_extend).  The longest cached prefix wins.
    const CachedStructure* base = nullptr;
    for (const CachedStructure& e : C.entries) {
        if (e.n_vars != n_vars || e.cons.size() >= cons.size() || e.cons.empty() || e.has_var_ids != (var_ids != nullptr)) continue;
        if (base && e.cons.size() <= base->cons.size()) continue;
        if (std::memcmp(e.cons.data(), cons.data(), e.cons.size() * sizeof(ezpz_constraint_t)) != 0) continue;
        if (var_ids && std::memcmp(e.var_ids.data(), var_ids, n_vars * sizeof(uint32_t)) != 0) continue;
        base = &e;
    }
    int32_t rc;
    if (base) {
        ++C.extended;
        rc = ezpz_b200_structure_extend(base->S, cons.data() + base->cons.size(), (uint32_t)(cons.size() - base->cons.size()), out, detail);
    } else {
        rc = ezpz_b200_structure_create(cons.data(), (uint32_t)cons.size(), var_ids, n_vars, out, detail);
    }
    if (rc != EZPZ_OK) return rc;
    auto held = [&] {
        uint64_t v = 0;
        for (const CachedStructure& e : C.entries) v += e.n_vars;
        return v;
    };
    while (!C.entries.empty() && (C.entries.size() >= kCacheEntries || held() + n_vars > kCacheVarBudget)) {
        size_t oldest = 0;
        for (size_t k = 1; k < C.entries.size(); ++k)
            if (C.entries[k].last_use < C.entries[oldest].last_use) oldest = k;
        ezpz_b200_structure_destroy(C.entries[oldest].S);
        C.entries.erase(C.entries.begin() + oldest);
    }
    CachedStructure e;
    e.hash = h;
    e.cons = cons;
    if (var_ids) e.var_ids.assign(var_ids, var_ids + n_vars);
    e.has_var_ids = var_ids != nullptr;
    e.n_vars = n_vars;
    e.S = *out;
    e.last_use = ++C.clock;
    C.entries.push_back(std::move(e));
    *owned = false;
    return EZPZ_OK;
}

constexpr double kEpsilon = 1e-4;  // lib.rs:43
bool nearly_eq(double a, double b) { return std::fabs(a - b) < kEpsilon; }  // warnings.rs:85-87

struct Warn {
    int64_t about;
    uint32_t kind;
    uint32_t count;
    double angle;
};

struct Level {
    std::vector<double> finals;
    std::vector<uint64_t> unsat;
    std::vector<uint32_t> under;
    std::vector<Warn> warnings;
    uint64_t iterations = 0;
    bool converged = false;
    uint32_t priority = 0;
    uint32_t num_vars = 0, num_eqs = 0;
    int32_t path = 0;
};

void export_level(const Level& L, ezpz_outcome_t* out) {
    if (out->final_values && !L.finals.empty()) std::memcpy(out->final_values, L.finals.data(), L.finals.size() * sizeof(double));
    out->n_unsatisfied = (uint32_t)L.unsat.size();
    if (out->unsatisfied) std::copy(L.unsat.begin(), L.unsat.end(), out->unsatisfied);
    out->n_underconstrained = (uint32_t)L.under.size();
    if (out->underconstrained) std::copy(L.under.begin(), L.under.end(), out->underconstrained);
    out->n_warnings = (uint32_t)L.warnings.size();
    if (out->warnings)
        for (uint32_t k = 0; k < std::min<uint32_t>(out->warnings_cap, out->n_warnings); ++k) {
            out->warnings[k].about_constraint = L.warnings[k].about;
            out->warnings[k].kind = L.warnings[k].kind;
            out->warnings[k].count = L.warnings[k].count;
            out->warnings[k].angle_deg = L.warnings[k].angle;
        }
    out->iterations = L.iterations;
    out->converged = L.converged ? 1 : 0;
    out->priority_solved = L.priority;
    out->num_vars = L.num_vars;
    out->num_eqs = L.num_eqs;
    out->path_used = L.path;
}

// solve_inner (lib.rs:265-356) for one subset.  `ids[k]` = original request index of subset entry k.
int32_t solve_level(ezpz_context_t* ctx, const std::vector<ezpz_constraint_t>& cons, const std::vector<uint64_t>& ids,
                    const std::vector<double>& angles, uint32_t level_priority, const uint32_t* var_ids,
                    const double* guesses, uint32_t n_vars, const ezpz_config_t* config, bool analysis, Level& L,
                    ezpz_error_detail_t* detail) {
    const uint32_t nc = (uint32_t)cons.size();
    L = Level();
    L.num_vars = n_vars;
    for (const auto& c : cons) L.num_eqs += ezk::kKinds[c.kind].rows;
    L.priority = level_priority;
    // warnings::lint — about_constraint is the ORIGINAL request index (constraint.id)
    for (uint32_t k = 0; k < nc; ++k) {
        if (cons[k].kind != EZPZ_K_LINES_AT_ANGLE || cons[k].flags != EZPZ_ANGLE_OTHER) continue;
        const double deg = angles.empty() ? NAN : angles[k];
        if (std::isnan(deg)) continue;
        if (nearly_eq(deg, 0.0) || nearly_eq(deg, 360.0) || nearly_eq(deg, 180.0)) L.warnings.push_back({(int64_t)ids[k], 1u, 1u, deg});
        else if (nearly_eq(deg, 90.0) || nearly_eq(deg, -90.0)) L.warnings.push_back({(int64_t)ids[k], 2u, 1u, deg});
    }
    ezpz_structure_t* S = nullptr;
    bool owned = true;
    int32_t rc = cached_structure(ctx, cons, var_ids, n_vars, &S, &owned, detail);
    if (rc != EZPZ_OK) {
        if (rc == EZPZ_ERR_MISSING_GUESS && detail) detail->constraint_id = ids[detail->constraint_id];
        return rc;
    }
    uint64_t nnz = 0;
    ezpz_b200_structure_dims(S, nullptr, nullptr, &nnz, nullptr, nullptr, nullptr);
    L.finals.assign(n_vars, 0.0);
    std::vector<uint32_t> unsat((nc + 31) / 32, 0), degen(nc, 0);
    std::vector<double> jac(analysis ? nnz : 0);
    uint32_t iterations = 0;
    uint8_t status = 0;
    ezpz_one_io_t io;
    std::memset(&io, 0, sizeof io);
    io.guesses = guesses;
    io.final_values = L.finals.data();
    io.iterations = &iterations;
    io.status = &status;
    io.unsat_mask = unsat.data();
    io.degen_count = degen.data();
    // structures of the batched kernel: the freedom analysis rides in the same call on the Jacobian the kernel leaves on the
    // device (ezpz_batch_io_t::under_mask); large systems export the Jacobian and are analysed after the solve
    int32_t path = -1;
    ezpz_b200_structure_ordering(S, &path, nullptr, nullptr, nullptr, nullptr, nullptr);
    const bool fused = analysis && path == 0;
    std::vector<uint32_t> under_mask((n_vars + 31) / 32, 0);
    if (fused) {
        ezpz_batch_io_t b;
        std::memset(&b, 0, sizeof b);
        b.guesses = guesses;
        b.final_values = L.finals.data();
        b.iterations = &iterations;
        b.status = &status;
        b.unsat_mask = unsat.data();
        b.degen_count = degen.data();
        b.under_mask = under_mask.data();
        L.path = 0;
        rc = ezpz_b200_solve_batch(ctx, S, config, 1, &b, detail);
    } else {
        io.jacobian = analysis ? jac.data() : nullptr;
        io.path_used = &L.path;
        rc = ezpz_b200_solve_one(ctx, S, config, &io, detail);
    }
    if (rc == EZPZ_OK) {
        L.iterations = iterations;
        L.converged = (status & EZPZ_ST_CONVERGED) != 0;
        // Degenerate warnings: about_constraint is the index inside the current subset (solver.rs:327,343)
        for (uint32_t k = 0; k < nc; ++k)
            if (degen[k]) L.warnings.push_back({(int64_t)k, 0u, degen[k], NAN});
        for (uint32_t k = 0; k < nc; ++k)
            if (unsat[k >> 5] & (1u << (k & 31u))) L.unsat.push_back(ids[k]);
        if (status & EZPZ_ST_SOLVE_ERROR) rc = EZPZ_ERR_SOLVE;
    }
    if (rc == EZPZ_OK && analysis) {
        if (!fused) rc = ezpz_b200_freedom_analysis(ctx, S, 1, jac.data(), under_mask.data(), detail);
        if (rc == EZPZ_OK)
            for (uint32_t j = 0; j < n_vars; ++j)
                if (under_mask[j >> 5] & (1u << (j & 31u))) L.under.push_back(j);
    }
    if (owned) ezpz_b200_structure_destroy(S);
    return rc;
}

}  // namespace

extern "C" void ezpz_b200_context_clear_cache(ezpz_context_t* ctx) { ezs::release_structure_cache(ctx); }

namespace ezs {
void release_structure_cache(ezpz_context* ctx) {
    if (!ctx || !ctx->structure_cache) return;
    StructureCache* C = static_cast<StructureCache*>(ctx->structure_cache);
    for (CachedStructure& e : C->entries) ezpz_b200_structure_destroy(e.S);
    delete C;
    ctx->structure_cache = nullptr;
}
}  // namespace ezs

extern "C" int32_t ezpz_b200_solve(ezpz_context_t* ctx, const ezpz_constraint_t* cons, const uint32_t* priorities,
                                   const double* angles_deg, uint32_t n_cons, const uint32_t* var_ids,
                                   const double* guesses, uint32_t n_vars, const ezpz_config_t* config,
                                   int32_t analysis, ezpz_outcome_t* outcome, ezpz_error_detail_t* detail) {
    if (!outcome || !config || (n_vars && !guesses) || (n_cons && !cons)) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    outcome->n_unsatisfied = outcome->n_underconstrained = outcome->n_warnings = 0;
    outcome->iterations = 0;
    outcome->converged = 0;
    outcome->priority_solved = 0;
    outcome->num_vars = n_vars;
    outcome->num_eqs = 0;
    outcome->path_used = -1;
    for (uint32_t k = 0; k < n_cons; ++k)
        if (cons[k].kind >= EZPZ_K_COUNT) return EZPZ_ERR_INVALID_ARGUMENT;
    if (n_cons == 0) {  // lib.rs:155-170
        if (outcome->final_values && n_vars) std::memcpy(outcome->final_values, guesses, n_vars * sizeof(double));
        outcome->converged = 1;
        return EZPZ_OK;
    }
    if (!ctx) return EZPZ_ERR_NO_DEVICE;
    std::vector<uint32_t> levels;
    for (uint32_t k = 0; k < n_cons; ++k) levels.push_back(priorities ? priorities[k] : 0u);
    std::sort(levels.begin(), levels.end());
    levels.erase(std::unique(levels.begin(), levels.end()), levels.end());
    // Constraint::set_from_initial_values reads guesses BY ID (lib.rs:172-186).  With ids 0..n-1 the device
    // resolves Undefined sides itself; with an explicit id list that is not the identity, resolve here so
    // that the by-id semantics is kept.
    std::vector<ezpz_constraint_t> all(cons, cons + n_cons);
    // canonical records: whatever a caller left in the ids a kind does not use, or in p0 / p1 of a kind without a scalar,
    // must not reach the kernels' tables nor split the topology cache into entries that differ only in garbage
    for (auto& c : all) {
        const ezk::KindInfo& ki = ezk::kKinds[c.kind];
        for (uint32_t k = ki.n_ids; k < 8; ++k) c.ids[k] = 0;
        const bool angle = c.kind == EZPZ_K_LINES_AT_ANGLE || c.kind == EZPZ_K_ARC_ANGLE || c.kind == EZPZ_K_POINTS_AT_ANGLE;
        const bool scalar = c.kind == EZPZ_K_DISTANCE || c.kind == EZPZ_K_VERTICAL_DISTANCE || c.kind == EZPZ_K_HORIZONTAL_DISTANCE ||
                            c.kind == EZPZ_K_FIXED || c.kind == EZPZ_K_CIRCLE_RADIUS || c.kind == EZPZ_K_ARC_RADIUS ||
                            c.kind == EZPZ_K_POINT_LINE_DISTANCE || c.kind == EZPZ_K_VERTICAL_POINT_LINE_DISTANCE ||
                            c.kind == EZPZ_K_HORIZONTAL_POINT_LINE_DISTANCE || c.kind == EZPZ_K_ARC_LENGTH;
        if (!angle) c.p1 = 0.0;
        if (!angle && !scalar) c.p0 = 0.0;
    }
    bool identity = true;
    if (var_ids)
        for (uint32_t k = 0; k < n_vars; ++k) identity = identity && var_ids[k] == k;
    if (!identity) {
        uint32_t max_id = 0;
        for (uint32_t k = 0; k < n_vars; ++k) max_id = std::max(max_id, var_ids[k]);
        std::vector<double> by_id((size_t)max_id + 1, 0.0);
        for (uint32_t k = 0; k < n_vars; ++k) by_id[var_ids[k]] = guesses[k];
        auto val = [&](uint32_t id) { return id < by_id.size() ? by_id[id] : 0.0; };
        for (auto& c : all) {
            if (c.flags != EZPZ_SIDE_UNDEFINED) continue;
            if (c.kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE) {
                const double ux = val(c.ids[2]) - val(c.ids[0]), uy = val(c.ids[3]) - val(c.ids[1]);
                const double vx = val(c.ids[4]) - val(c.ids[0]), vy = val(c.ids[5]) - val(c.ids[1]);
                c.flags = (ux * vy - uy * vx >= 0.0) ? EZPZ_LINE_SIDE_LEFT : EZPZ_LINE_SIDE_RIGHT;
            } else if (c.kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE) {
                const double dist = ezpz_b200_hypot(val(c.ids[0]) - val(c.ids[3]), val(c.ids[1]) - val(c.ids[4]));
                const double ar = val(c.ids[2]), br = val(c.ids[5]);
                const double r_int = std::fabs(std::fabs(ar - br) - dist), r_ext = std::fabs(ar + br - dist);
                c.flags = (r_int < r_ext) ? EZPZ_CIRCLE_SIDE_INTERIOR : EZPZ_CIRCLE_SIDE_EXTERIOR;
            }
        }
    }
    bool have = false;
    Level best;
    for (uint32_t level : levels) {
        std::vector<ezpz_constraint_t> subset;
        std::vector<uint64_t> ids;
        std::vector<double> angles;
        uint32_t lowest = 0;
        for (uint32_t k = 0; k < n_cons; ++k) {
            const uint32_t p = priorities ? priorities[k] : 0u;
            if (p <= level) {
                subset.push_back(all[k]);
                ids.push_back(k);
                if (angles_deg) angles.push_back(angles_deg[k]);
                lowest = std::max(lowest, p);
            }
        }
        Level cur;
        const int32_t rc = solve_level(ctx, subset, ids, angles, lowest, var_ids, guesses, n_vars, config, analysis != 0, cur, detail);
        if (rc != EZPZ_OK) {  // lib.rs:239-244: fall back to the previous level, else report the error
            if (have) {
                export_level(best, outcome);
                return EZPZ_OK;
            }
            export_level(cur, outcome);
            return rc;
        }
        if (!cur.unsat.empty()) {  // lib.rs:228-234
            export_level(have ? best : cur, outcome);
            return EZPZ_OK;
        }
        best = std::move(cur);
        have = true;
    }
    export_level(best, outcome);
    return EZPZ_OK;
}


// The same priority loop for a batch of problems of one topology: one structure and one batched solve per level.
extern "C" int32_t ezpz_b200_solve_batch_priorities(ezpz_context_t* ctx, const ezpz_constraint_t* cons, const uint32_t* priorities,
                                                    uint32_t n_cons, uint32_t n_vars, const ezpz_config_t* config, uint64_t batch,
                                                    const double* guesses, const double* params, double* final_values,
                                                    uint32_t* iterations, uint8_t* status, uint32_t* priority_solved,
                                                    uint32_t* unsat_mask, ezpz_error_detail_t* detail) {
    if (!config || (n_cons && !cons) || !final_values || !iterations || !status || (batch && n_vars && !guesses))
        return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    for (uint32_t k = 0; k < n_cons; ++k)
        if (cons[k].kind >= EZPZ_K_COUNT) return EZPZ_ERR_INVALID_ARGUMENT;
    const size_t uw = (n_cons + 31) / 32;
    if (n_cons == 0) {  // lib.rs:155-170
        if (batch && n_vars) std::memcpy(final_values, guesses, (size_t)batch * n_vars * sizeof(double));
        for (uint64_t b = 0; b < batch; ++b) {
            iterations[b] = 0;
            status[b] = EZPZ_ST_CONVERGED;
            if (priority_solved) priority_solved[b] = 0;
        }
        return EZPZ_OK;
    }
    if (!ctx) return EZPZ_ERR_NO_DEVICE;
    if (batch == 0) return EZPZ_OK;
    std::vector<uint32_t> levels;
    for (uint32_t k = 0; k < n_cons; ++k) levels.push_back(priorities ? priorities[k] : 0u);
    std::sort(levels.begin(), levels.end());
    levels.erase(std::unique(levels.begin(), levels.end()), levels.end());
    // per problem: 0 = no level kept yet, 1 = a satisfied level is kept, 2 = stopped
    std::vector<uint8_t> state(batch, 0);
    std::vector<uint64_t> active(batch);
    for (uint64_t b = 0; b < batch; ++b) active[b] = b;
    std::vector<double> g_sub, p_sub, f_sub;
    std::vector<uint32_t> it_sub, un_sub;
    std::vector<uint8_t> st_sub;
    for (size_t li = 0; li < levels.size() && !active.empty(); ++li) {
        const uint32_t level = levels[li];
        std::vector<ezpz_constraint_t> subset;
        std::vector<uint32_t> ids;
        uint32_t lowest = 0;
        for (uint32_t k = 0; k < n_cons; ++k) {
            const uint32_t p = priorities ? priorities[k] : 0u;
            if (p <= level) {
                subset.push_back(cons[k]);
                ids.push_back(k);
                lowest = std::max(lowest, p);
            }
        }
        const uint32_t nc = (uint32_t)subset.size();
        const size_t uws = (nc + 31) / 32;
        ezpz_structure_t* S = nullptr;
        int32_t rc = ezpz_b200_structure_create(subset.data(), nc, nullptr, n_vars, &S, detail);
        if (rc != EZPZ_OK) {  // lib.rs:239-244: problems that already keep a level keep it; otherwise the error is the answer
            if (rc == EZPZ_ERR_MISSING_GUESS && detail) detail->constraint_id = ids[detail->constraint_id];
            return li == 0 ? rc : EZPZ_OK;
        }
        // the still-undecided problems, compacted
        const uint64_t nb = active.size();
        g_sub.resize((size_t)nb * n_vars);
        f_sub.resize((size_t)nb * n_vars);
        it_sub.resize(nb);
        st_sub.resize(nb);
        un_sub.assign((size_t)nb * uws, 0u);
        if (params) p_sub.resize((size_t)nb * nc);
        for (uint64_t q = 0; q < nb; ++q) {
            std::memcpy(g_sub.data() + q * n_vars, guesses + active[q] * n_vars, n_vars * sizeof(double));
            if (params)
                for (uint32_t k = 0; k < nc; ++k) p_sub[q * nc + k] = params[active[q] * n_cons + ids[k]];
        }
        ezpz_batch_io_t io;
        std::memset(&io, 0, sizeof io);
        io.guesses = g_sub.data();
        io.params = params ? p_sub.data() : nullptr;
        io.final_values = f_sub.data();
        io.iterations = it_sub.data();
        io.status = st_sub.data();
        io.unsat_mask = un_sub.data();
        rc = ezpz_b200_solve_batch(ctx, S, config, nb, &io, detail);
        ezpz_b200_structure_destroy(S);
        if (rc != EZPZ_OK) return li == 0 ? rc : EZPZ_OK;
        std::vector<uint64_t> next;
        for (uint64_t q = 0; q < nb; ++q) {
            const uint64_t b = active[q];
            const bool unsat = (st_sub[q] & EZPZ_ST_UNSATISFIED) != 0;
            const bool keep = !unsat || state[b] == 0;  // an unsatisfied level is kept only when nothing else is (lib.rs:232-234)
            if (keep) {
                std::memcpy(final_values + b * n_vars, f_sub.data() + q * n_vars, n_vars * sizeof(double));
                iterations[b] = it_sub[q];
                status[b] = st_sub[q];
                if (priority_solved) priority_solved[b] = lowest;
                if (unsat_mask) {
                    uint32_t* dst = unsat_mask + b * uw;
                    std::fill(dst, dst + uw, 0u);
                    for (uint32_t k = 0; k < nc; ++k)
                        if (un_sub[q * uws + (k >> 5)] & (1u << (k & 31u))) dst[ids[k] >> 5] |= 1u << (ids[k] & 31u);
                }
            }
            if (unsat) state[b] = 2;
            else {
                state[b] = 1;
                next.push_back(b);
            }
        }
        active.swap(next);
    }
    return EZPZ_OK;
}
