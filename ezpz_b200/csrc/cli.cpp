// cli.cpp — `ezpz-b200`, the twin of ezpz-cli (ezpz-cli/src/main.rs:47-170) on the GPU path.
//   ezpz-b200 -f <problem.md | -> [--show-points]
// Parses the text problem, solves it once, then re-solves NUM_ITERS_BENCHMARK = 100 times and prints the same
// lines the reference CLI prints (warnings, unsatisfied constraints, "Problem size: R rows, V vars",
// "Iterations needed", "Solved up to priority", "Solved in N us (mean over 100 iterations)", "i.e. N solves
// per second", and with --show-points the points/circles/arcs to two decimals).  PNG output (-o) is out of
// scope.  Like the reference, every re-solve repeats the whole solve call including the
// structure analysis (main.rs:95-98).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/ezpz_b200.h"

static const char* kKindNames[] = {"LineTangentToCircle", "CircleTangentToCircle", "Distance", "DistanceVar", "VerticalDistance",
                                   "HorizontalDistance", "Vertical", "Horizontal", "LinesAtAngle", "Fixed", "ScalarEqual",
                                   "PointsCoincident", "CircleRadius", "LinesEqualLength", "ArcRadius", "Arc", "Midpoint",
                                   "PointLineDistance", "VerticalPointLineDistance", "HorizontalPointLineDistance", "Symmetric",
                                   "PointArcCoincident", "ArcLength", "ArcAngle", "PointsAtAngle"};
static const uint8_t kRows[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 2, 1, 2, 1, 1, 1, 2, 2, 2, 1, 2};

int main(int argc, char** argv) {
    std::string path;
    bool show_points = false;
    for (int i = 1; i < argc; ++i) {
        if ((!std::strcmp(argv[i], "-f") || !std::strcmp(argv[i], "--filepath")) && i + 1 < argc) path = argv[++i];
        else if (!std::strcmp(argv[i], "--show-points")) show_points = true;
        else if (!std::strcmp(argv[i], "-o") || !std::strcmp(argv[i], "--image-path")) {
            std::fprintf(stderr, "Error: image output is not part of the GPU path\n");
            return 1;
        }
    }
    if (path.empty()) {
        std::fprintf(stderr, "usage: ezpz-b200 -f <file|-> [--show-points]\n");
        return 2;
    }
    std::stringstream buf;
    if (path == "-") buf << std::cin.rdbuf();
    else {
        std::ifstream f(path);
        if (!f) {
            std::fprintf(stderr, "Error: cannot read %s\n", path.c_str());
            return 1;
        }
        buf << f.rdbuf();
    }
    const std::string text = buf.str();
    ezpz_error_detail_t det;
    ezpz_problem_t* problem = nullptr;
    int32_t rc = ezpz_b200_problem_parse(text.data(), text.size(), &problem, &det);
    if (rc != EZPZ_OK) {
        std::fprintf(stderr, "Error: %s\n", det.message);
        return 1;
    }
    const ezpz_constraint_t* cons = nullptr;
    const double* guesses = nullptr;
    uint32_t n_cons = 0, n_vars = 0;
    rc = ezpz_b200_problem_system(problem, &cons, &n_cons, &guesses, &n_vars, &det);
    if (rc != EZPZ_OK) {
        std::fprintf(stderr, "Error: %s\n", det.message);
        return 1;
    }
    const double* angles = nullptr;
    ezpz_b200_problem_angles_deg(problem, &angles);
    ezpz_context_t* ctx = nullptr;
    rc = ezpz_b200_context_create(0, &ctx, &det);
    if (rc != EZPZ_OK) {
        std::fprintf(stderr, "Error: %s (%s)\n", ezpz_b200_status_name(rc), det.message);
        return 1;
    }
    ezpz_config_t cfg;
    ezpz_b200_config_default(&cfg);
    std::vector<double> finals(n_vars ? n_vars : 1);
    std::vector<uint64_t> unsat(n_cons ? n_cons : 1);
    std::vector<ezpz_warning_t> warns(2 * n_cons + 8);
    ezpz_outcome_t out;
    auto solve = [&]() {
        std::memset(&out, 0, sizeof out);
        out.final_values = finals.data();
        out.unsatisfied = unsat.data();
        out.warnings = warns.data();
        out.warnings_cap = (uint32_t)warns.size();
        return ezpz_b200_solve(ctx, cons, nullptr, angles, n_cons, nullptr, guesses, n_vars, &cfg, 0, &out, &det);
    };
    const auto t0 = std::chrono::steady_clock::now();
    rc = solve();
    auto print_size = [&](uint32_t eqs, uint32_t vars) { std::printf("Problem size: %u rows, %u vars\n", eqs, vars); };
    if (rc != EZPZ_OK) {
        uint32_t eqs = 0;
        for (uint32_t c = 0; c < n_cons; ++c) eqs += kRows[cons[c].kind];
        print_size(eqs, n_vars);
        std::fprintf(stderr, "Could not solve system: %s %s\n", ezpz_b200_status_name(rc), det.message);
        std::fprintf(stderr, eqs > n_vars ? "Your system might be overconstrained. Try removing constraints.\n"
                                          : "You might have contradictory constraints.\n");
        return 1;
    }
    const ezpz_outcome_t first = out;
    std::vector<double> first_finals = finals;
    std::vector<uint64_t> first_unsat(unsat.begin(), unsat.begin() + first.n_unsatisfied);
    std::vector<ezpz_warning_t> first_warns(warns.begin(), warns.begin() + std::min<uint32_t>(first.n_warnings, (uint32_t)warns.size()));
    const int kIters = 100;  // NUM_ITERS_BENCHMARK
    for (int k = 0; k < kIters; ++k) solve();
    const auto elapsed = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    const long long per_iter = elapsed / kIters;
    if (!first_warns.empty()) {
        std::printf("Warnings:\n");
        for (const auto& w : first_warns) {
            if (w.kind == 0)
                std::printf("\tThis geometry is degenerate, meaning two points are so close together that they practically "
                            "overlap. This is probably unintentional, you probably should place your initial guesses further "
                            "apart or choose different constraints.\n");
            else if (w.kind == 1) std::printf("\tInstead of constraining to %gdeg, constrain to Parallel\n", w.angle_deg);
            else std::printf("\tInstead of constraining to %gdeg, constraint to Perpendicular\n", w.angle_deg);
        }
    }
    if (!first_unsat.empty()) {
        std::printf("Not all constraints were satisfied:\n");
        for (uint64_t c : first_unsat) std::printf("\t%llu: %s\n", (unsigned long long)c, kKindNames[cons[c].kind]);
    }
    print_size(first.num_eqs, first.num_vars);
    std::printf("Iterations needed: %llu\n", (unsigned long long)first.iterations);
    std::printf("Solved up to priority: %u\n", first.priority_solved);
    if (!first.converged) std::printf("Error: solver did not converge!\n");
    std::printf("Solved in %lld\xce\xbcs (mean over %d iterations)\n", per_iter, kIters);
    std::printf("i.e. %lld solves per second\n", per_iter > 0 ? 1000000LL / per_iter : 0LL);
    if (show_points) {
        const uint32_t np = ezpz_b200_problem_count(problem, 0), ncirc = ezpz_b200_problem_count(problem, 1),
                       narc = ezpz_b200_problem_count(problem, 2);
        std::printf("Points:\n");
        for (uint32_t i = 0; i < np; ++i)
            std::printf("\t%s: (%.2f, %.2f)\n", ezpz_b200_problem_label(problem, 0, i), first_finals[2 * i], first_finals[2 * i + 1]);
        if (ncirc) {
            std::printf("Circles:\n");
            for (uint32_t i = 0; i < ncirc; ++i) {
                const double* v = &first_finals[2 * np + 3 * i];
                std::printf("\t%s: center = (%.2f, %.2f), radius = %.2f\n", ezpz_b200_problem_label(problem, 1, i), v[0], v[1], v[2]);
            }
        }
        if (narc) {
            std::printf("Arcs:\n");
            for (uint32_t i = 0; i < narc; ++i) {
                const double* v = &first_finals[2 * np + 3 * ncirc + 6 * i];
                std::printf("\t%s: center = (%.2f, %.2f), a = (%.2f, %.2f), b = (%.2f, %.2f)\n", ezpz_b200_problem_label(problem, 2, i),
                            v[4], v[5], v[0], v[1], v[2], v[3]);
            }
        }
    }
    ezpz_b200_context_destroy(ctx);
    ezpz_b200_problem_destroy(problem);
    return 0;
}
