"""Host logic, CPU only: the C ABI loads and exports every declared symbol, the text pipeline, the
structure analysis (pattern parity with the oracle, bit-exact), sharding."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ezpz_b200 as ez
import orc
import workloads as wl
from ezpz_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "ezpz_b200.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(ezpz_b200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    lib = C.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/ezpz_b200.h but not exported"
    assert sorted(native.SYMBOL_NAMES) == declared, "native.py must bind exactly the header's functions"
    assert native.lib().ezpz_b200_abi_version() == 4
    assert C.sizeof(native.Constraint) == 64


def test_rust_binding_declares_exactly_the_header():
    """integration/gpu.rs (the extern "C" block a maintainer adds to the crate) and include/ezpz_b200.h name the same functions,
    and the structs that cross the boundary have the same field lists."""
    with open(os.path.join(ROOT, "include", "ezpz_b200.h")) as f:
        header = f.read()
    with open(os.path.join(ROOT, "integration", "gpu.rs")) as f:
        rust = f.read()
    declared = sorted(set(re.findall(r"\b(ezpz_b200_[a-z0-9_]+)\s*\(", header)))
    bound = sorted(set(re.findall(r"pub fn (ezpz_b200_[a-z0-9_]+)\s*\(", rust)))
    assert bound == declared
    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (name, name), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return [re.sub(r"\[.*\]", "", d.split()[-1].lstrip("*")) for stmt in body.split(";") if stmt.strip()
                for d in [stmt.strip()] if d]
    def rs_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rust, re.S).group(1)
        return re.findall(r"pub ([a-z0-9_]+):", body)
    assert c_fields("ezpz_batch_io") == rs_fields("EzpzBatchIo")
    assert c_fields("ezpz_one_io") == rs_fields("EzpzOneIo")
    assert c_fields("ezpz_config") == rs_fields("EzpzConfig")
    assert c_fields("ezpz_outcome") == rs_fields("EzpzOutcome")
    assert "EZPZ_B200_ABI_VERSION: u32 = %d" % native.lib().ezpz_b200_abi_version() in rust


def test_config_default():
    cfg = native.Config()
    native.lib().ezpz_b200_config_default(C.byref(cfg))
    assert (cfg.max_iterations, cfg.residual_tolerance, cfg.step_tolerance, cfg.initial_lambda) == (35, 1e-8, 1e-12, 1e-9)


def test_tiny_worked_pattern():
    """SURVEY.md §8c worked example (test_cases/tiny)."""
    recs, n, g, _ = wl.system_from_text(wl.fixture_text("tiny"))
    pat = ez.Structure(recs, n).pattern()
    assert list(pat["csc_col_ptr"]) == [0, 2, 3, 4, 5] and list(pat["csc_row_idx"]) == [0, 3, 1, 3, 2]
    assert list(pat["csr_row_ptr"]) == [0, 1, 2, 3, 5] and list(pat["csr_col_idx"]) == [0, 1, 3, 0, 2]


@pytest.mark.parametrize("name", sorted(wl.fixtures()))
def test_patterns_match_oracle_bit_exactly(name):
    recs, n, g, _ = wl.system_from_text(wl.fixture_text(name))
    st = ez.Structure(recs, n)
    pat = st.pattern()
    rc, op = orc.pattern(recs, n)
    assert rc == 0
    for k in ("csc_col_ptr", "csc_row_idx", "csr_row_ptr", "csr_col_idx", "cons_row0"):
        assert np.array_equal(pat[k], op[k]), k
    assert (pat["m"], pat["nnz"]) == (op["m"], op["nnz"])
    # pattern of A and of the Cholesky factor: product is CSC with the diagonal first, oracle is by rows
    pa, oc = st.pattern_a(), orc.pattern_chol(recs, n)
    def cols_to_rows(col_ptr, row_idx, strict):
        rows = [[] for _ in range(n)]
        for j in range(n):
            for p in range(col_ptr[j], col_ptr[j + 1]):
                if not (strict and row_idx[p] == j):
                    rows[row_idx[p]].append(j)
        return rows
    a_rows = cols_to_rows(pa["a_col_ptr"], pa["a_row_idx"], False)
    l_rows = cols_to_rows(pa["l_col_ptr"], pa["l_row_idx"], True)
    for i in range(n):
        assert a_rows[i] == list(oc["a_col"][oc["a_row_ptr"][i]:oc["a_row_ptr"][i + 1]])
        assert l_rows[i] == list(oc["l_col"][oc["l_row_ptr"][i]:oc["l_row_ptr"][i + 1]])


def test_patterns_random_and_massive():
    from test_gpu_parity import random_constraints
    rng = np.random.default_rng(3)
    for _ in range(10):
        recs = ez.records(random_constraints(rng, 80, 40))
        pat = ez.Structure(recs, 40).pattern()
        rc, op = orc.pattern(recs, 40)
        for k in ("csc_col_ptr", "csc_row_idx", "csr_row_ptr", "csr_col_idx"):
            assert np.array_equal(pat[k], op[k])
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500))
    st = ez.Structure(recs, n)
    assert (st.m, st.n, st.nnz, st.n_components) == (2000, 2000, 2500, 1500)
    rc, op = orc.pattern(recs, n)
    assert np.array_equal(st.pattern()["csr_col_idx"], op["csr_col_idx"])
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(600))
    st = ez.Structure(recs, n)
    assert (st.m, st.n, st.nnz) == (2400, 2400, 3000)  # the checked-in problem.md


def test_structure_errors():
    Cn = ez.Constraint
    with pytest.raises(ez.EzpzError) as e:
        ez.Structure(ez.records([Cn.Fixed(0, 0.0), Cn.Fixed(7, 1.0)]), 4)
    assert e.value.name == "MissingGuess" and e.value.constraint_id == 1 and e.value.variable == 7
    with pytest.raises(ez.EzpzError) as e:  # id is among the guesses but beyond the matrix (FaerMatrix)
        ez.Structure(ez.records([Cn.Fixed(9, 0.0)]), 2, var_ids=[3, 9])
    assert e.value.name == "FaerMatrix"
    recs = ez.records([Cn.Fixed(0, 0.0)])
    recs[0]["kind"] = 99
    with pytest.raises(ez.EzpzError) as e:
        ez.Structure(recs, 1)
    assert e.value.name == "InvalidArgument"
    # PointsCoincident with only x ids guessed: first missing id comes from row 1 (solver.rs:448-478)
    with pytest.raises(ez.EzpzError) as e:
        ez.Structure(ez.records([Cn.PointsCoincident(ez.DatumPoint.new_xy(0, 1), ez.DatumPoint.new_xy(2, 3))]), 2,
                     var_ids=[0, 2])
    assert e.value.name == "MissingGuess" and e.value.variable == 1


def test_text_errors_and_quirks():
    T = ez.textual
    with pytest.raises(T.TextualError) as e:
        T.Problem("# constraints\npoint p\n\n# guesses\nq roughly (0, 0)\n").to_constraint_system()
    assert e.value.name == "TextMissingGuess"
    with pytest.raises(T.TextualError) as e:
        T.Problem("# constraints\npoint p\n\n# guesses\np roughly (0, 0)\nghost roughly (1, 1)\n").to_constraint_system()
    assert e.value.name == "TextUnusedGuesses" and "ghost" in e.value.message
    with pytest.raises(T.TextualError) as e:
        T.Problem("# constraints\npoint p\nmissing.x = 2.5\n\n# guesses\np roughly (0, 0)\n").to_constraint_system()
    assert e.value.name == "TextUndefinedPoint"
    for bad in ["", "# constraints\n", "# constraints\npoint p\n# guesses\np roughly (0,0)\n",
                "# constraints\npoint p\n\n# guesses\np roughly (0,0)\n\n\nextra"]:
        with pytest.raises(T.TextualError) as e:
            T.Problem(bad)
        assert e.value.name == "Parse"
    # sqrt(...) number expressions, deg/rad angles, no spaces around '='
    cs = T.Problem("# constraints\npoint a\npoint b\npoint c\npoint d\ndistance(a, b, sqrt(sqrt(16)))\n"
                   "lines_at_angle(a, b, c, d, 90deg)\na.x=1e0\n\n# guesses\na roughly (0,0)\nb roughly (1, 0)\n"
                   "c roughly (-.5,+2.)\nd roughly (3,3)").to_constraint_system()
    assert cs.constraints["p0"][0] == 2.0 and cs.angles_deg[1] == 90.0 and cs.constraints["p0"][2] == 1.0
    assert list(cs.initial_guesses[4:6]) == [-0.5, 2.0]
    s, c = ez.angle_sincos(90.0 * (np.pi / 180.0))
    assert (cs.constraints["p0"][1], cs.constraints["p1"][1]) == (c, s)
    # `A.center = (x, y)` on an ARC is silently dropped (executor.rs:273-283); circles keep it
    cs = T.Problem("# constraints\narc a\na.center = (1, 2)\nis_arc(a)\n\n# guesses\na.center roughly (0,0)\n"
                   "a.a roughly (1,0)\na.b roughly (0,1)\n").to_constraint_system()
    assert len(cs.constraints) == 1 and cs.constraints["kind"][0] == ez.K_ARC
    assert list(cs.constraints["ids"][0][:6]) == [0, 1, 2, 3, 4, 5]
    cs = T.Problem("# constraints\ncircle c\nc.center = (1, 2)\nradius(c, 3)\n\n# guesses\nc.center roughly (0,0)\n"
                   "c.radius roughly 1\n").to_constraint_system()
    assert list(cs.constraints["kind"]) == [ez.K_FIXED, ez.K_FIXED, ez.K_CIRCLE_RADIUS]
    assert [int(r["ids"][0]) for r in cs.constraints[:2]] == [0, 1] and int(cs.constraints["ids"][2][2]) == 2


def test_text_readers_agree():
    """The product's parser/executor twin (textual.cpp) and the oracle-side reader (tests/textual_twin.py, independent, pure
    Python) must produce the same records, guesses and label lists on every fixture and generated file: the parity tests feed
    the oracle through the latter, so a label-resolution or numbering bug in the product cannot hide behind GPU == oracle."""
    import textual_twin
    texts = [wl.fixture_text(name) for name in sorted(wl.fixtures())]
    texts += [wl.massive_problem_text(60, False), wl.massive_problem_text(25, True)]
    # every instruction form the fixtures do not use together: circle + arc in one file (the arc-base quirk), sqrt(), rad
    texts.append("# constraints\npoint p\npoint q\npoint r\npoint s\ncircle c\narc a\nline(p, q)\nradius(c, sqrt(4))\n"
                 "tangent(p, q, c)\nc.center = (1, 2)\nc.center.x = 1\na.center.y = 0\nis_arc(a)\narc_radius(a, 2)\n"
                 "arc_length(a, 1.5)\npoint_arc_coincident(p, a)\ncoincident(q, r)\nmidpoint(p, q, r)\nsymmetric(p, q, r, s)\n"
                 "lines_at_angle(p, q, r, s, 0.5rad)\nlines_at_angle(p, q, r, s, 33deg)\npoint_line_distance(s, p, q, 2.5)\n"
                 "parallel(p, q, r, s)\nperpendicular(p, q, r, s)\nlines_equal_length(p, q, r, s)\nhorizontal(p, q)\n"
                 "vertical(p, s)\ndistance(p, q, 3)\np.x = 0\n\n"
                 "# guesses\np roughly (0, 0)\nq roughly (1, 1)\nr roughly (3, 1)\ns roughly (4, -1)\n"
                 "c.center roughly (2, 2)\nc.radius roughly 1\na.center roughly (5, 5)\na.a roughly (6, 5)\na.b roughly (5, 6)\n")
    for text in texts:
        pr, pn, pg, pcs = wl.product_system_from_text(text)
        tw = textual_twin.parse(text)
        assert pr.tobytes() == tw.constraints.tobytes()
        assert np.array_equal(pg, tw.initial_guesses)
        assert (pcs.inner_points, pcs.inner_circles, pcs.inner_arcs) == (tw.inner_points, tw.inner_circles, tw.inner_arcs)
        a, b = pcs.angles_deg, tw.angles_deg
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-15)


def test_variable_numbering_points_circles_arcs():
    """executor.rs:41-108: points first, then circles (cx,cy,r), then arcs (a,b,center)."""
    cs = ez.textual.Problem(wl.fixture_text("circle_tangent")).to_constraint_system()
    assert cs.num_vars == 7 and cs.inner_points == ["p", "q"] and cs.inner_circles == ["a"]
    tan = cs.constraints[cs.constraints["kind"] == ez.K_LINE_TANGENT_TO_CIRCLE][0]
    assert list(tan["ids"][:7]) == [0, 1, 2, 3, 4, 5, 6] and tan["flags"] == 0
    cs = ez.textual.Problem(wl.fixture_text("parc_coincident")).to_constraint_system()
    pac = cs.constraints[cs.constraints["kind"] == ez.K_POINT_ARC_COINCIDENT][0]
    assert list(pac["ids"]) == [2, 3, 4, 5, 6, 7, 0, 1]
    assert list(cs.initial_guesses) == [4, 3, 0, 4, 4, 0, 0.1, 0.2]


def test_shard_range():
    L = native.lib()
    for batch, world in [(65536, 8), (10, 3), (5, 8), (0, 4), (1000003, 7)]:
        covered = []
        for rank in range(world):
            b, e = C.c_uint64(), C.c_uint64()
            L.ezpz_b200_shard_range(batch, rank, world, C.byref(b), C.byref(e))
            covered.append((b.value, e.value))
        assert covered[0][0] == 0 and covered[-1][1] == batch
        assert all(covered[k][1] == covered[k + 1][0] for k in range(world - 1))
        sizes = [e - b for b, e in covered]
        assert max(sizes) - min(sizes) <= 1


def test_host_math_matches_oracle_bitwise():
    """The product's own libm restatement (dmath.cuh, host side) against the oracle's, bit for bit."""
    rng = np.random.default_rng(2)
    L, O = native.lib(), orc.lib()
    for k in range(20000):
        a, b = rng.uniform(-1, 1, 2) * 10.0 ** rng.integers(-12, 12)
        assert L.ezpz_b200_hypot(a, b) == O.orc_fn_hypot(a, b)
        s, c = ez.angle_sincos(a)
        assert s == O.orc_fn_sin(a) and c == O.orc_fn_cos(a)
    for a in [0.0, -0.0, np.pi, np.pi / 2, -np.pi / 4, 1e-30, 0.7853981633974483, 0.7853981633974484, 1e5, -3e5]:
        s, c = ez.angle_sincos(a)
        assert s == O.orc_fn_sin(a) and c == O.orc_fn_cos(a)


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product path must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ez.EzpzError) as e:
        ez.Context(0)
    assert e.value.name in ("NoDevice", "Cuda")
    with pytest.raises(Exception):
        ez.solve([ez.ConstraintRequest.highest_priority(ez.Constraint.Fixed(0, 0.0))], [(0, 1.0)])
    # the one case that needs no device: an empty request list returns the guesses (lib.rs:155-170)
    out = ez.solve([], [(0, 0.5)])
    assert list(out.final_values()) == [0.5] and out.iterations() == 0 and out.converged()


def test_large_system_ordering_host_analysis():
    """sparse_direct.cpp on the CPU: the elimination order is a permutation, natural for the block-diagonal
    massive_parallel_system (shallow tree), nested dissection with a logarithmic tree for the chain sketch; the oracle
    solves to the same answer (identical iteration count, 1e-9) whichever order it is given."""
    import orc
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500, False))
    od = ez.Structure(recs, n).ordering()
    assert od["path"] == 1 and not od["nested"] and od["n_levels"] <= 4 and od["sum_chunk"] == 64
    assert np.array_equal(od["elim_order"], np.arange(n))
    heights = {}
    for cells in (64, 1024, 4096):
        recs, n, g, exact = wl.chain_sketch(cells)
        st = ez.Structure(recs, n)
        od = st.ordering()
        assert od["path"] == 1 and od["nested"]
        assert sorted(od["elim_order"].tolist()) == list(range(n))
        assert od["nnz_l"] < 32 * n  # panel storage (explicit zeros of the relaxed supernodes included) stays linear in n
        heights[cells] = od["n_levels"]
        if cells <= 1024:
            a = orc.solve_inner(recs, g)
            b = orc.solve_inner_ordered(recs, g, od["elim_order"], od["sum_chunk"])
            assert a.iterations == b.iterations and a.converged and b.converged
            assert np.abs(a.final_values - b.final_values).max() < 1e-9
            c = orc.solve_inner_ordered(recs, g, None, 0)
            assert np.array_equal(a.final_values, c.final_values)
    # 64x more cells add a few stages to the supernode tree, they do not multiply it
    assert heights[4096] <= heights[64] + 12 and heights[4096] < 40
    # a small system takes the batched kernel: no large programme
    recs, n, g, _ = wl.system_from_text(wl.fixture_text("square"))
    assert ez.Structure(recs, n).ordering()["path"] == 0


def _check_role_program(st, roles):
    """Invariants of build_role_blob (structure.cpp), checked exactly: every op of the sequential tape appears once, each
    role's ops keep the sequential order, and between two barriers no role writes a slot another role reads or writes."""
    seq = st.role_program(1)["roles"][0]["ops"]
    assert not any("barrier" in o for o in seq)
    rp = st.role_program(roles)
    key = lambda o: (o["dst"], o["shape"], o["added"], o["fin"], tuple(o["pairs"]))
    # the sequential tape writes a destination several times (A, then L); tell the ops apart by their rank among equals
    rank, seen = {}, {}
    for k, o in enumerate(seq):
        seen[key(o)] = seen.get(key(o), 0) + 1
        rank[(key(o), seen[key(o)])] = k
    covered = []
    n_bar = None
    for role in rp["roles"]:
        counts, last = {}, -1
        bars = 0
        for o in role["ops"]:
            if "barrier" in o:
                bars += 1
                continue
            counts[key(o)] = counts.get(key(o), 0) + 1
            k = rank[(key(o), counts[key(o)])] if (key(o), counts[key(o)]) in rank else None
            assert k is not None, "op not in the sequential tape"
            covered.append(k)
        assert n_bar in (None, bars), "every role must hold the same number of barriers"
        n_bar = bars
    assert n_bar == rp["barriers"]
    # (equal ops executed by different roles may swap ranks; as a multiset the coverage must be exact)
    assert sorted(covered) == list(range(len(seq)))
    # per-role order: positions of a role's ops in the sequential tape must be increasing when ops are matched greedily
    for role in rp["roles"]:
        pos, used = 0, set()
        for o in role["ops"]:
            if "barrier" in o:
                continue
            while pos < len(seq) and (pos in used or key(seq[pos]) != key(o)):
                pos += 1
            assert pos < len(seq), "a role's ops are not a subsequence of the sequential tape"
            used.add(pos)
    # hazards inside an epoch
    epochs = [[] for _ in range(n_bar + 1)]
    for r, role in enumerate(rp["roles"]):
        e = 0
        for o in role["ops"]:
            if "barrier" in o:
                e += 1
                continue
            reads = {s for p in o["pairs"] for s in p}
            if o["shape"] == 3:  # TAPE_BACKWARD starts from V[dst]
                reads.add(o["dst"])
            if o["shape"] != 2:  # every shape but TAPE_PIVOT multiplies by V[fin]
                reads.add(o["fin"])
            epochs[e].append((r, reads, o["dst"]))
    for e, ops in enumerate(epochs):
        writes = {}
        for r, reads, dst in ops:
            writes.setdefault(dst, set()).add(r)
        for dst, rs in writes.items():
            assert len(rs) == 1, f"epoch {e}: slot {dst} written by roles {rs}"
        for r, reads, dst in ops:
            for s in reads:
                assert writes.get(s, {r}) == {r}, f"epoch {e}: role {r} reads slot {s} written by {writes[s]}"
    # constraint lists partition the constraints; the split ranges tile x, r and J
    cons = sorted(c for role in rp["roles"] for c in role["constraints"])
    assert cons == list(range(st.n_cons))
    for what, total in (("x", st.n), ("r", st.m), ("j", st.nnz)):
        edges = [role[what] for role in rp["roles"]]
        assert edges[0][0] == 0 and edges[-1][1] == total and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    return rp


@pytest.mark.parametrize("roles", [2, 3, 4])
def test_role_programs_of_the_batched_kernel(roles):
    """The batched kernel splits a problem's work over cooperating warps (device.cu); the split must be a pure
    re-distribution of the sequential tape with a barrier in front of every cross-role dependency."""
    import ezpz_b200 as ez
    import workloads as wl
    names = [n for n in sorted(wl.fixtures())]
    for name in names:
        recs, n, g, _ = wl.system_from_text(wl.fixture_text(name))
        st = ez.Structure(recs, n)
        rp = _check_role_program(st, roles)
        if name == "two_rectangles" and roles == 2:
            # two independent 8-variable blocks: each role assembles, factorises and solves one of them, no barrier at all
            assert rp["barriers"] == 0
            pairs = [sum(len(o.get("pairs", [])) for o in role["ops"]) for role in rp["roles"]]
            assert pairs[0] == pairs[1]
    from test_gpu_parity import random_constraints
    rng = np.random.default_rng(31 + roles)
    checked = 0
    for trial in range(8):
        cons = random_constraints(rng, 26 + 2 * trial, 20)[-(8 + 3 * trial):]  # (the first 25 are one of each kind)
        st = ez.Structure(ez.records(cons), 20)
        try:
            _check_role_program(st, roles)
            checked += 1
        except ez.EzpzError as e:  # too large for the thread-per-problem kernel
            assert e.name == "Unsupported"
    assert checked >= 5


def test_threaded_host_analysis_equals_sequential(monkeypatch):
    """every phase of the host analysis that runs on host threads for large systems (pattern of J as an
    atomic bucket pass, scatter slots, pattern of A, adjacency, symbolic factorisation over independent subtrees of the
    elimination tree, supernode panels, update lists, products of A, tile order) produces exactly what one thread does:
    the fingerprint hashes every array the device reads."""
    cases = [wl.chain_sketch(20000)[:2], wl.grid_truss(120)[:2]]
    recs, n = wl.system_from_text(wl.massive_problem_text(500, False))[:2]  # natural-order system: 80 disjoint copies
    copies = []
    for k in range(80):
        c = recs.copy()
        c["ids"] += np.uint32(k * n)
        copies.append(c)
    cases.append((np.concatenate(copies), 80 * n))
    for recs, n in cases:
        fps = {}
        for threads in ("1", "3", "16"):
            monkeypatch.setenv("EZPZ_B200_HOST_THREADS", threads)
            st = ez.Structure(recs, n)
            fps[threads] = (st.fingerprint(), st.ordering()["path"])
        assert fps["1"] == fps["3"] == fps["16"], fps
        assert fps["1"][1] == 1


def test_structure_extend_equals_analysis_of_the_longer_list():
    """ezpz_b200_structure_extend (add constraints to an analysed sketch, tests.rs:748-897): for systems of the batched
    kernel and for natural-order systems the result is the structure ezpz_b200_structure_create gives for the concatenated
    list, array for array; errors name the constraint by its index in the concatenated list."""
    for name in sorted(wl.fixtures()):
        recs, n, g, _ = wl.system_from_text(wl.fixture_text(name))
        if len(recs) < 2:
            continue
        for cut in {1, len(recs) // 2, len(recs) - 1}:
            ext = ez.Structure(recs[:cut], n).extend(recs[cut:])
            full = ez.Structure(recs, n)
            assert ext.fingerprint() == full.fingerprint(), (name, cut)
            assert (ext.m, ext.n, ext.nnz, ext.n_cons) == (full.m, full.n, full.nnz, full.n_cons)
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500, False))
    assert ez.Structure(recs[:1000], n).extend(recs[1000:]).fingerprint() == ez.Structure(recs, n).fingerprint()
    assert ez.Structure(recs, n).extend(recs[:0]).fingerprint() == ez.Structure(recs, n).fingerprint()
    Cn = ez.Constraint
    base = ez.Structure(ez.records([Cn.Fixed(0, 0.0), Cn.Fixed(2, 1.0)]), 3, var_ids=[0, 2, 5])
    with pytest.raises(ez.EzpzError) as e:  # id 1 is not among the guesses the base was created with
        base.extend(ez.records([Cn.Fixed(0, 0.5), Cn.Fixed(1, 1.0)]))
    assert e.value.name == "MissingGuess" and e.value.constraint_id == 3 and e.value.variable == 1
    with pytest.raises(ez.EzpzError) as e:  # among the guesses but beyond the matrix
        base.extend(ez.records([Cn.Fixed(5, 0.5)]))
    assert e.value.name == "FaerMatrix"


def test_structure_extend_keeps_the_elimination_order(monkeypatch):
    """Large-path systems: the extended structure keeps the base's elimination order (no second dissection), its patterns
    are those of the longer list, the factor stays about as sparse as a fresh analysis makes it, and the oracle solves to
    the same answer in that order (identical iterations, 1e-9) as in the natural one."""
    import orc
    for recs, n, g, exact in (wl.chain_sketch(304), wl.grid_truss(20)):
        k = len(recs) - 9
        base = ez.Structure(recs[:k], n)
        ext, full = base.extend(recs[k:]), ez.Structure(recs, n)
        ob, oe, of = base.ordering(), ext.ordering(), full.ordering()
        assert ob["path"] == oe["path"] == of["path"] == 1 and ob["nested"] and oe["nested"]
        assert np.array_equal(oe["elim_order"], ob["elim_order"])
        pe, pf = ext.pattern(), full.pattern()
        for key in pe:
            assert np.array_equal(pe[key], pf[key]), key
        assert oe["nnz_l"] <= 1.25 * of["nnz_l"] and oe["sum_chunk"] == of["sum_chunk"]
        a = orc.solve_inner(recs, g)
        b = orc.solve_inner_ordered(recs, g, oe["elim_order"], oe["sum_chunk"])
        assert a.iterations == b.iterations and a.converged == b.converged
        assert np.abs(a.final_values - b.final_values).max() < 1e-9
        # constraints that couple no new pair of variables (here: repeats of existing ones) leave A's pattern alone: the whole
        # sparse-direct schedule is taken over from the base and only the product lists of A = JtJ are rebuilt; the result is
        # what the full re-derivation in the kept order gives
        again = np.concatenate([recs[3:5], recs[k:k + 2]])
        fast = full.extend(again)
        monkeypatch.setenv("EZPZ_B200_EXTEND_FULL", "1")
        slow = full.extend(again)
        monkeypatch.delenv("EZPZ_B200_EXTEND_FULL")
        assert fast.fingerprint() == slow.fingerprint()
        assert np.array_equal(fast.ordering()["elim_order"], of["elim_order"]) and fast.m == full.m + slow.m - full.m > full.m


def test_structure_extend_random_lists():
    """Random constraint lists of every kind, cut at random points: the extended structure has the patterns and row numbering of
    the whole list; where no elimination order is carried over (batched-kernel and natural-order systems) it IS the fresh
    analysis, array for array; where one is, it is a permutation and the schedule re-derived in that order (EZPZ_B200_EXTEND_FULL)
    is the schedule taken over when A kept its pattern."""
    from test_gpu_parity import random_constraints
    rng = np.random.default_rng(11)
    kept = 0
    for trial in range(40):
        n = int(rng.integers(6, 120))
        recs = ez.records(random_constraints(rng, int(rng.integers(4, 3 * n)), n))
        cut = int(rng.integers(1, len(recs)))
        base, full = ez.Structure(recs[:cut], n), ez.Structure(recs, n)
        ext = base.extend(recs[cut:])
        pe, pf = ext.pattern(), full.pattern()
        for key in pe:
            assert np.array_equal(pe[key], pf[key]), (trial, key)
        ob, oe = base.ordering(), ext.ordering()
        if ob["path"] == 1 and oe["path"] == 1:
            kept += 1
            assert np.array_equal(oe["elim_order"], ob["elim_order"]) and sorted(oe["elim_order"].tolist()) == list(range(n))
            if not ob["nested"]:
                assert ext.fingerprint() == full.fingerprint(), trial
        else:
            assert ext.fingerprint() == full.fingerprint(), trial
        again = recs[rng.integers(0, len(recs), 3)]
        fast = full.extend(again)
        os.environ["EZPZ_B200_EXTEND_FULL"] = "1"
        try:
            slow = full.extend(again)
        finally:
            del os.environ["EZPZ_B200_EXTEND_FULL"]
        assert fast.fingerprint() == slow.fingerprint(), trial
    assert kept >= 5


@pytest.mark.filterwarnings("ignore::DeprecationWarning")  # (Python's generic warning about fork() in a threaded process)
def test_host_pool_survives_fork():
    """The host analysis runs its phases on a persistent pool of threads (host_parallel.h).  A forked child inherits the pool
    object but none of its threads: it must start a pool of its own instead of waiting for workers that do not exist."""
    import multiprocessing as mp
    recs, n = wl.chain_sketch(20000)[:2]
    fp = ez.Structure(recs, n).fingerprint()  # (the parent's pool exists from here on)

    def child(q):
        q.put(ez.Structure(recs, n).fingerprint())

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=child, args=(q,))
    p.start()
    got = q.get(timeout=120)
    p.join(timeout=30)
    assert got == fp and p.exitcode == 0
    assert ez.Structure(recs, n).fingerprint() == fp  # and the parent's pool still works
