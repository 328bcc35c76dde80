// textual.cpp — the ezpz text problem format: parser and "executor" (text -> constraint records and
// initial guesses).  Host only; this is the call surface ezpz-cli reaches the solver through
// (ezpz-cli/src/main.rs:81-98: Problem::from_str -> to_constraint_system -> solve).
//
// Grammar and alternative order follow ezpz/src/textual/parser.rs:29-555 (a winnow combinator parser;
// here a hand-written backtracking recursive-descent parser with the same acceptance set); the
// instruction -> Constraint mapping and the variable numbering follow
// ezpz/src/textual/executor.rs:40-445 and geometry_variables.rs:11-177 (points 2i,2i+1; then circles
// cx,cy,r; then arcs a.x,a.y,b.x,b.y,c.x,c.y).  Two reference quirks are kept on purpose and
// documented in DESIGN.md: `arc_ids` ignores circles when computing the first arc id
// (geometry_variables.rs:92), and `A.center = (x, y)` for an ARC is silently dropped
// (executor.rs:273-283).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ezpz_b200.h"
#include "dmath.cuh"

namespace {

enum class Ins {
    DeclarePoint, DeclareCircle, DeclareArc, FixPointComponent, FixCenterPointComponent, Horizontal, Vertical,
    PointsCoincident, PointArcCoincident, Midpoint, Symmetric, Distance, Parallel, Perpendicular, AngleLine,
    CircleRadius, Tangent, ArcRadius, ArcLength, IsArc, PointLineDistance, Line, LinesEqualLength
};

struct Instruction {
    Ins op;
    std::string l[4];   // labels in textual order
    int component = 0;  // 0 = x, 1 = y
    double value = 0.0; // number argument
    bool degrees = false;
};

struct PointGuess {
    std::string label;
    double x, y;
};
struct ScalarGuess {
    std::string label;
    double v;
};

struct Cursor {
    const char* s;
    size_t n;
    size_t i = 0;
    bool eof() const { return i >= n; }
    char peek() const { return i < n ? s[i] : '\0'; }
};

bool lit(Cursor& c, const char* w) {
    size_t k = std::strlen(w);
    if (c.n - c.i < k || std::memcmp(c.s + c.i, w, k) != 0) return false;
    c.i += k;
    return true;
}
bool ch(Cursor& c, char x) {
    if (c.peek() != x || c.eof()) return false;
    ++c.i;
    return true;
}
void space0(Cursor& c) {
    while (!c.eof() && (c.s[c.i] == ' ' || c.s[c.i] == '\t')) ++c.i;
}
bool newline(Cursor& c) { return ch(c, '\n'); }
bool is_alnum(char x) { return (x >= '0' && x <= '9') || (x >= 'a' && x <= 'z') || (x >= 'A' && x <= 'Z'); }

bool label(Cursor& c, std::string& out) {  // parser.rs:495-499
    size_t b = c.i;
    while (!c.eof() && is_alnum(c.s[c.i])) ++c.i;
    if (c.i == b) return false;
    out.assign(c.s + b, c.i - b);
    return true;
}
bool label_opt_suffix(Cursor& c, std::string& out) {  // parser.rs:501-509
    if (!label(c, out)) return false;
    Cursor save = c;
    std::string suf;
    if (ch(c, '.') && label(c, suf)) {
        out.push_back('.');
        out += suf;
    } else {
        c = save;
    }
    return true;
}

bool ieq(const char* a, const char* b, size_t k) {
    for (size_t q = 0; q < k; ++q)
        if (std::tolower((unsigned char)a[q]) != b[q]) return false;
    return true;
}

// winnow::ascii::float, then digit1 (parser.rs:537-547)
bool number(Cursor& c, double& out) {
    size_t b = c.i, p = c.i;
    if (p < c.n && (c.s[p] == '+' || c.s[p] == '-')) ++p;
    size_t d0 = p;
    while (p < c.n && std::isdigit((unsigned char)c.s[p])) ++p;
    bool have_int = p > d0, have_frac = false;
    if (p < c.n && c.s[p] == '.') {
        size_t q = p + 1;
        while (q < c.n && std::isdigit((unsigned char)c.s[q])) ++q;
        if (q > p + 1 || have_int) {
            have_frac = q > p + 1;
            p = q;
        }
    }
    if (have_int || have_frac) {
        if (p < c.n && (c.s[p] == 'e' || c.s[p] == 'E')) {
            size_t q = p + 1;
            if (q < c.n && (c.s[q] == '+' || c.s[q] == '-')) ++q;
            size_t e0 = q;
            while (q < c.n && std::isdigit((unsigned char)c.s[q])) ++q;
            if (q > e0) p = q;
        }
        std::string tok(c.s + b, p - b);
        out = std::strtod(tok.c_str(), nullptr);
        c.i = p;
        return true;
    }
    // nan / inf / infinity, case-insensitive, optional sign
    size_t rem = c.n - d0;
    const bool neg = d0 > b && c.s[b] == '-';
    if (rem >= 8 && ieq(c.s + d0, "infinity", 8)) {
        c.i = d0 + 8;
        out = neg ? -INFINITY : INFINITY;
        return true;
    }
    if (rem >= 3 && ieq(c.s + d0, "inf", 3)) {
        c.i = d0 + 3;
        out = neg ? -INFINITY : INFINITY;
        return true;
    }
    if (rem >= 3 && ieq(c.s + d0, "nan", 3)) {
        c.i = d0 + 3;
        out = std::numeric_limits<double>::quiet_NaN();
        return true;
    }
    return false;
}

bool number_expr(Cursor& c, double& out) {  // parser.rs:549-555
    Cursor save = c;
    if (number(c, out)) return true;
    c = save;
    double inner;
    if (lit(c, "sqrt(") && number_expr(c, inner) && ch(c, ')')) {
        out = std::sqrt(inner);
        return true;
    }
    c = save;
    return false;
}

bool commasep(Cursor& c) {  // parser.rs:223-228
    space0(c);
    if (!ch(c, ',')) return false;
    space0(c);
    return true;
}

bool point_tuple(Cursor& c, double& x, double& y) {  // parser.rs:511-516: "(" ws num "," space0 num ")"
    Cursor save = c;
    if (ch(c, '(')) {
        space0(c);
        if (number(c, x) && ch(c, ',')) {
            space0(c);
            if (number(c, y) && ch(c, ')')) return true;
        }
    }
    c = save;
    return false;
}

// "(" ws L ("," L){k-1} ws ")"   (two_points / three_points / four_points inside_brackets)
bool labels_in_brackets(Cursor& c, int k, std::string* out) {
    if (!ch(c, '(')) return false;
    space0(c);
    for (int q = 0; q < k; ++q) {
        if (q && !commasep(c)) return false;
        if (!label(c, out[q])) return false;
    }
    space0(c);
    return ch(c, ')');
}

// keyword ws "(" ws L{k} [ "," numexpr|num|angle ] ")"
enum class Tail { None, Number, NumberExpr, Angle };
bool call(Cursor& c, const char* kw, int k, Tail tail, Instruction& ins) {
    Cursor save = c;
    if (!lit(c, kw)) return false;
    space0(c);
    bool ok = false;
    if (tail == Tail::None) {
        ok = labels_in_brackets(c, k, ins.l);
    } else if (ch(c, '(')) {
        space0(c);
        ok = true;
        for (int q = 0; q < k && ok; ++q) {
            if (q && !commasep(c)) ok = false;
            if (ok && !label(c, ins.l[q])) ok = false;
        }
        if (ok && k == 3 && tail == Tail::Number) {
            // three_labels_num (parser.rs:373-384): commasep, number, ws
            ok = commasep(c) && number(c, ins.value);
            if (ok) space0(c);
        } else if (ok && (k == 2 || k == 4)) {
            // (two_points|four_points, commasep, tail): the point lists eat trailing blanks first
            space0(c);
            ok = commasep(c);
            if (ok && tail == Tail::NumberExpr) ok = number_expr(c, ins.value);
            else if (ok && tail == Tail::Angle) {
                ok = number(c, ins.value);
                if (ok) {
                    if (lit(c, "deg")) ins.degrees = true;
                    else if (lit(c, "rad")) ins.degrees = false;
                    else ok = false;
                }
            } else if (ok && tail == Tail::Number) ok = number(c, ins.value);
        } else if (ok && k == 1) {
            // (parse_label, commasep, number | number_expr)
            ok = commasep(c);
            if (ok) ok = (tail == Tail::NumberExpr) ? number_expr(c, ins.value) : number(c, ins.value);
        }
        if (ok) ok = ch(c, ')');
    }
    if (!ok) c = save;
    return ok;
}

bool component(Cursor& c, int& comp) {
    if (ch(c, 'x')) {
        comp = 0;
        return true;
    }
    if (ch(c, 'y')) {
        comp = 1;
        return true;
    }
    return false;
}

bool eq_number(Cursor& c, double& v) {  // delimited(space0, '=', space0), parse_number
    space0(c);
    if (!ch(c, '=')) return false;
    space0(c);
    return number(c, v);
}

// One instruction line -> 1 or 2 instructions (parser.rs:386-442).
bool instruction(Cursor& c, std::vector<Instruction>& out) {
    space0(c);
    const Cursor start = c;
    Instruction ins;
    auto declare = [&](const char* kw, Ins op) {
        c = start;
        if (lit(c, kw)) {
            space0(c);
            if (label(c, ins.l[0])) {
                ins.op = op;
                out.push_back(ins);
                return true;
            }
        }
        return false;
    };
    if (declare("point", Ins::DeclarePoint)) return true;
    if (declare("circle", Ins::DeclareCircle)) return true;
    if (declare("arc", Ins::DeclareArc)) return true;
    c = start;  // L.x = v
    if (label(c, ins.l[0]) && ch(c, '.') && component(c, ins.component) && eq_number(c, ins.value)) {
        ins.op = Ins::FixPointComponent;
        out.push_back(ins);
        return true;
    }
    c = start;  // L.center.x = v
    if (label(c, ins.l[0]) && lit(c, ".center.") && component(c, ins.component) && eq_number(c, ins.value)) {
        ins.op = Ins::FixCenterPointComponent;
        out.push_back(ins);
        return true;
    }
    c = start;  // L[.L] = (a, b)  -> two FixPointComponent
    {
        double px, py;
        if (label_opt_suffix(c, ins.l[0])) {
            space0(c);
            if (ch(c, '=')) {
                space0(c);
                if (point_tuple(c, px, py)) {
                    ins.op = Ins::FixPointComponent;
                    ins.component = 0;
                    ins.value = px;
                    out.push_back(ins);
                    ins.component = 1;
                    ins.value = py;
                    out.push_back(ins);
                    return true;
                }
            }
        }
    }
    struct Form {
        const char* kw;
        int k;
        Tail tail;
        Ins op;
    };
    static const Form forms[] = {
        {"horizontal", 2, Tail::None, Ins::Horizontal},
        {"coincident", 2, Tail::None, Ins::PointsCoincident},
        {"point_arc_coincident", 2, Tail::None, Ins::PointArcCoincident},
        {"midpoint", 3, Tail::None, Ins::Midpoint},
        {"symmetric", 4, Tail::None, Ins::Symmetric},
        {"vertical", 2, Tail::None, Ins::Vertical},
        {"distance", 2, Tail::NumberExpr, Ins::Distance},
        {"parallel", 4, Tail::None, Ins::Parallel},
        {"perpendicular", 4, Tail::None, Ins::Perpendicular},
        {"lines_at_angle", 4, Tail::Angle, Ins::AngleLine},
        {"radius", 1, Tail::NumberExpr, Ins::CircleRadius},
        {"tangent", 3, Tail::None, Ins::Tangent},
        {"arc_radius", 1, Tail::Number, Ins::ArcRadius},
        {"arc_length", 1, Tail::Number, Ins::ArcLength},
        {"is_arc", 1, Tail::None, Ins::IsArc},
        {"point_line_distance", 3, Tail::Number, Ins::PointLineDistance},
        {"line", 2, Tail::None, Ins::Line},
        {"lines_equal_length", 4, Tail::None, Ins::LinesEqualLength},
    };
    for (const Form& f : forms) {
        c = start;
        Instruction q;
        q.op = f.op;
        if (call(c, f.kw, f.k, f.tail, q)) {
            out.push_back(q);
            return true;
        }
    }
    c = start;
    return false;
}

bool guess(Cursor& c, std::vector<PointGuess>& pts, std::vector<ScalarGuess>& scs) {  // parser.rs:85-128
    const Cursor start = c;
    space0(c);
    std::string lab;
    if (!label_opt_suffix(c, lab)) {
        c = start;
        return false;
    }
    space0(c);
    if (!lit(c, "roughly")) {
        c = start;
        return false;
    }
    space0(c);
    double x, y;
    if (point_tuple(c, x, y)) {
        pts.push_back({lab, x, y});
        return true;
    }
    if (number(c, x)) {
        scs.push_back({lab, x});
        return true;
    }
    c = start;
    return false;
}

}  // namespace

struct ezpz_problem {
    std::vector<Instruction> instructions;
    std::vector<std::string> points, circles, arcs;
    std::vector<std::pair<std::string, std::string>> lines;
    std::vector<PointGuess> point_guesses;
    std::vector<ScalarGuess> scalar_guesses;
    // filled by ezpz_b200_problem_system
    bool built = false;
    std::vector<ezpz_constraint_t> cons;
    std::vector<double> guesses;
    std::vector<double> angles_deg;
};

namespace {

void fail_at(ezpz_error_detail_t* d, const Cursor& c, const char* what) {
    if (!d) return;
    size_t line = 1, col = 1;
    for (size_t q = 0; q < c.i && q < c.n; ++q) {
        if (c.s[q] == '\n') {
            ++line;
            col = 1;
        } else ++col;
    }
    std::snprintf(d->message, sizeof d->message, "parse error at line %zu column %zu: expected %s", line, col, what);
    d->a = line;
    d->b = col;
}

ezpz_constraint_t blank(uint32_t kind) {
    ezpz_constraint_t r;
    std::memset(&r, 0, sizeof r);
    r.kind = kind;
    r.weight = 1.0;
    return r;
}

struct Pt {
    uint32_t x, y;
};

}  // namespace

extern "C" {

int32_t ezpz_b200_problem_parse(const char* text, uint64_t len, ezpz_problem_t** out, ezpz_error_detail_t* detail) {
    if (!text || !out) return EZPZ_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (detail) std::memset(detail, 0, sizeof *detail);
    Cursor c{text, (size_t)len};
    ezpz_problem* P = new (std::nothrow) ezpz_problem();
    if (!P) return EZPZ_ERR_INVALID_ARGUMENT;
    auto bail = [&](const char* what) {
        fail_at(detail, c, what);
        delete P;
        return (int32_t)EZPZ_ERR_PARSE;
    };
    // "# constraints\n"
    if (!ch(c, '#')) return bail("'# constraints' header");
    space0(c);
    if (!lit(c, "constraints") || !newline(c)) return bail("'# constraints' header");
    // separated(1.., instruction, newline)
    if (!instruction(c, P->instructions)) return bail("an instruction");
    for (;;) {
        Cursor save = c;
        if (!newline(c)) break;
        if (!instruction(c, P->instructions)) {
            c = save;
            break;
        }
    }
    if (!newline(c) || !newline(c)) return bail("a blank line before '# guesses'");
    space0(c);
    if (!ch(c, '#')) return bail("'# guesses' header");
    space0(c);
    if (!lit(c, "guesses") || !newline(c)) return bail("'# guesses' header");
    if (!guess(c, P->point_guesses, P->scalar_guesses)) return bail("a guess");
    for (;;) {
        Cursor save = c;
        if (!newline(c)) break;
        if (!guess(c, P->point_guesses, P->scalar_guesses)) {
            c = save;
            break;
        }
    }
    newline(c);  // opt(newline)
    space0(c);
    if (!c.eof()) return bail("end of input");
    for (const Instruction& ins : P->instructions) {
        if (ins.op == Ins::DeclarePoint) P->points.push_back(ins.l[0]);
        if (ins.op == Ins::DeclareCircle) P->circles.push_back(ins.l[0]);
        if (ins.op == Ins::DeclareArc) P->arcs.push_back(ins.l[0]);
        if (ins.op == Ins::Line) P->lines.emplace_back(ins.l[0], ins.l[1]);
    }
    *out = P;
    return EZPZ_OK;
}

void ezpz_b200_problem_destroy(ezpz_problem_t* p) { delete p; }

int32_t ezpz_b200_problem_system(ezpz_problem_t* P, const ezpz_constraint_t** cons, uint32_t* n_cons,
                                 const double** guesses, uint32_t* n_vars, ezpz_error_detail_t* detail) {
    if (!P) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    auto text_err = [&](int32_t code, const char* fmt, const std::string& lab) {
        if (detail) std::snprintf(detail->message, sizeof detail->message, fmt, lab.c_str());
        return code;
    };
    if (!P->built) {
        P->cons.clear();
        P->guesses.clear();
        P->angles_deg.clear();
        // ---- guesses -> variables (executor.rs:41-117).  Later duplicates of a label overwrite earlier ones.
        std::unordered_map<std::string, std::pair<double, double>> gp;
        std::vector<std::string> gp_order;
        for (const PointGuess& g : P->point_guesses) {
            if (!gp.count(g.label)) gp_order.push_back(g.label);
            gp[g.label] = {g.x, g.y};
        }
        std::unordered_map<std::string, double> gs;
        std::vector<std::string> gs_order;
        for (const ScalarGuess& g : P->scalar_guesses) {
            if (!gs.count(g.label)) gs_order.push_back(g.label);
            gs[g.label] = g.v;
        }
        auto take_point = [&](const std::string& lab, std::pair<double, double>& v) {
            auto it = gp.find(lab);
            if (it == gp.end()) return false;
            v = it->second;
            gp.erase(it);
            return true;
        };
        std::pair<double, double> v;
        for (const std::string& p : P->points) {
            if (!take_point(p, v)) return text_err(EZPZ_ERR_TEXT_MISSING_GUESS, "No guess was given for point %s", p);
            P->guesses.push_back(v.first);
            P->guesses.push_back(v.second);
        }
        for (const std::string& cl : P->circles) {
            const std::string cen = cl + ".center", rad = cl + ".radius";
            if (!take_point(cen, v)) return text_err(EZPZ_ERR_TEXT_MISSING_GUESS, "No guess was given for point %s", cen);
            auto it = gs.find(rad);
            if (it == gs.end()) return text_err(EZPZ_ERR_TEXT_MISSING_GUESS, "No guess was given for point %s", rad);
            P->guesses.push_back(v.first);
            P->guesses.push_back(v.second);
            P->guesses.push_back(it->second);
            gs.erase(it);
        }
        for (const std::string& ar : P->arcs) {
            std::pair<double, double> ce, a, b;
            if (!take_point(ar + ".center", ce)) return text_err(EZPZ_ERR_TEXT_MISSING_GUESS, "No guess was given for point %s", ar + ".center");
            if (!take_point(ar + ".a", a)) return text_err(EZPZ_ERR_TEXT_MISSING_GUESS, "No guess was given for point %s", ar + ".a");
            if (!take_point(ar + ".b", b)) return text_err(EZPZ_ERR_TEXT_MISSING_GUESS, "No guess was given for point %s", ar + ".b");
            for (double q : {a.first, a.second, b.first, b.second, ce.first, ce.second}) P->guesses.push_back(q);
        }
        if (!gp.empty()) {
            for (const std::string& lab : gp_order)
                if (gp.count(lab)) return text_err(EZPZ_ERR_TEXT_UNUSED_GUESSES, "You gave a guess for points which weren't defined: %s", lab);
        }
        if (!gs.empty()) {
            for (const std::string& lab : gs_order)
                if (gs.count(lab)) return text_err(EZPZ_ERR_TEXT_UNUSED_GUESSES, "You gave a guess for points which weren't defined: %s", lab);
        }
        const uint32_t np = (uint32_t)P->points.size(), ncirc = (uint32_t)P->circles.size();
        // Label -> index tables (first declaration wins, as a linear scan would find it).  The reference resolves
        // every label with a linear scan and a format! per candidate (executor.rs:121-174), quadratic in the number
        // of points; here a lookup is one hash probe, so a 160,000-line file resolves in milliseconds.
        auto make_index = [](const std::vector<std::string>& v) {
            std::unordered_map<std::string, int> m;
            m.reserve(v.size() * 2);
            for (size_t q = 0; q < v.size(); ++q) m.emplace(v[q], (int)q);
            return m;
        };
        const std::unordered_map<std::string, int> ix_points = make_index(P->points), ix_circles = make_index(P->circles),
                                                   ix_arcs = make_index(P->arcs);
        auto index_of = [&](const std::vector<std::string>& v, const std::string& s) -> int {
            const auto& m = &v == &P->points ? ix_points : (&v == &P->circles ? ix_circles : ix_arcs);
            const auto it = m.find(s);
            return it == m.end() ? -1 : it->second;
        };
        auto strip = [](const std::string& s, const char* suffix, std::string& base) {
            size_t k = std::strlen(suffix);
            if (s.size() < k || s.compare(s.size() - k, k, suffix) != 0) return false;
            base = s.substr(0, s.size() - k);
            return true;
        };
        auto circle_base = [&](int ci) { return 2 * np + 3 * (uint32_t)ci; };
        // geometry_variables.rs:92: the first arc id is computed as 2*num_points, circles ignored (kept).
        auto arc_base = [&](int ai) { return 2 * np + 6 * (uint32_t)ai; };
        (void)ncirc;
        // datum_point_for_label (executor.rs:121-174)
        auto point_for = [&](const std::string& lab, Pt& out) -> bool {
            int k = index_of(P->points, lab);
            if (k >= 0) {
                out = {2u * (uint32_t)k, 2u * (uint32_t)k + 1};
                return true;
            }
            std::string base;
            if (strip(lab, ".center", base)) {
                if ((k = index_of(P->circles, base)) >= 0) {
                    out = {circle_base(k), circle_base(k) + 1};
                    return true;
                }
                if ((k = index_of(P->arcs, base)) >= 0) {
                    out = {arc_base(k) + 4, arc_base(k) + 5};
                    return true;
                }
            }
            if (strip(lab, ".a", base) && (k = index_of(P->arcs, base)) >= 0) {
                out = {arc_base(k), arc_base(k) + 1};
                return true;
            }
            if (strip(lab, ".b", base) && (k = index_of(P->arcs, base)) >= 0) {
                out = {arc_base(k) + 2, arc_base(k) + 3};
                return true;
            }
            return false;
        };
        auto undefined = [&](const std::string& lab) {
            return text_err(EZPZ_ERR_TEXT_UNDEFINED_POINT, "You referred to the point %s but it was never defined", lab);
        };
        auto push = [&](const ezpz_constraint_t& r, double angle_deg) {
            P->cons.push_back(r);
            P->angles_deg.push_back(angle_deg);
        };
        const double kNaN = std::numeric_limits<double>::quiet_NaN();
        auto set_pts = [](ezpz_constraint_t& r, int at, Pt p) {
            r.ids[at] = p.x;
            r.ids[at + 1] = p.y;
        };
        auto arc_of = [&](const std::string& lab, ezpz_constraint_t& r, std::string& bad) -> bool {
            Pt ce, a, b;  // center first, as the reference resolves it (executor.rs:209-213)
            if (!point_for(lab + ".center", ce)) { bad = lab + ".center"; return false; }
            if (!point_for(lab + ".a", a)) { bad = lab + ".a"; return false; }
            if (!point_for(lab + ".b", b)) { bad = lab + ".b"; return false; }
            set_pts(r, 0, a);
            set_pts(r, 2, b);
            set_pts(r, 4, ce);
            return true;
        };
        auto circle_of = [&](const std::string& lab, ezpz_constraint_t& r, int at, std::string& bad) -> bool {
            Pt ce;
            if (!point_for(lab + ".center", ce)) { bad = lab + ".center"; return false; }
            int k = index_of(P->circles, lab);  // datum_distance_for_label (executor.rs:175-187)
            if (k < 0) { bad = lab + ".radius"; return false; }
            set_pts(r, at, ce);
            r.ids[at + 2] = circle_base(k) + 2;
            return true;
        };
        for (const Instruction& ins : P->instructions) {
            std::string bad;
            Pt a, b, c2, d2;
            switch (ins.op) {
                case Ins::DeclarePoint:
                case Ins::DeclareCircle:
                case Ins::DeclareArc:
                case Ins::Line: break;
                case Ins::CircleRadius: {
                    ezpz_constraint_t r = blank(EZPZ_K_CIRCLE_RADIUS);
                    if (!circle_of(ins.l[0], r, 0, bad)) return undefined(bad);
                    r.p0 = ins.value;
                    push(r, kNaN);
                } break;
                case Ins::ArcRadius:
                case Ins::ArcLength:
                case Ins::IsArc: {
                    ezpz_constraint_t r = blank(ins.op == Ins::ArcRadius ? EZPZ_K_ARC_RADIUS
                                                : ins.op == Ins::ArcLength ? EZPZ_K_ARC_LENGTH : EZPZ_K_ARC);
                    if (!arc_of(ins.l[0], r, bad)) return undefined(bad);
                    r.p0 = ins.op == Ins::IsArc ? 0.0 : ins.value;
                    push(r, kNaN);
                } break;
                case Ins::PointLineDistance: {  // point, line_p0, line_p1 ; line resolved first (executor.rs:231-236)
                    if (!point_for(ins.l[1], b)) return undefined(ins.l[1]);
                    if (!point_for(ins.l[2], c2)) return undefined(ins.l[2]);
                    if (!point_for(ins.l[0], a)) return undefined(ins.l[0]);
                    ezpz_constraint_t r = blank(EZPZ_K_POINT_LINE_DISTANCE);
                    set_pts(r, 0, a);
                    set_pts(r, 2, b);
                    set_pts(r, 4, c2);
                    r.p0 = ins.value;
                    push(r, kNaN);
                } break;
                case Ins::Tangent: {  // tangent(line_p0, line_p1, circle)
                    ezpz_constraint_t r = blank(EZPZ_K_LINE_TANGENT_TO_CIRCLE);
                    if (!circle_of(ins.l[2], r, 4, bad)) return undefined(bad);
                    if (!point_for(ins.l[0], a)) return undefined(ins.l[0]);
                    if (!point_for(ins.l[1], b)) return undefined(ins.l[1]);
                    set_pts(r, 0, a);
                    set_pts(r, 2, b);
                    r.flags = EZPZ_SIDE_UNDEFINED;
                    push(r, kNaN);
                } break;
                case Ins::FixPointComponent: {  // executor.rs:259-289
                    int k = index_of(P->points, ins.l[0]);
                    std::string base;
                    if (k >= 0) {
                        ezpz_constraint_t r = blank(EZPZ_K_FIXED);
                        r.ids[0] = 2u * (uint32_t)k + (uint32_t)ins.component;
                        r.p0 = ins.value;
                        push(r, kNaN);
                    } else if (strip(ins.l[0], ".center", base)) {
                        int ci = index_of(P->circles, base);
                        if (ci >= 0) {
                            ezpz_constraint_t r = blank(EZPZ_K_FIXED);
                            r.ids[0] = circle_base(ci) + (uint32_t)ins.component;
                            r.p0 = ins.value;
                            push(r, kNaN);
                        }  // an arc's ".center = (..)" falls through silently, as in the reference
                    } else {
                        return undefined(ins.l[0]);
                    }
                } break;
                case Ins::FixCenterPointComponent: {  // executor.rs:290-320
                    int ci = index_of(P->circles, ins.l[0]);
                    int ai = index_of(P->arcs, ins.l[0]);
                    ezpz_constraint_t r = blank(EZPZ_K_FIXED);
                    if (ci >= 0) r.ids[0] = circle_base(ci) + (uint32_t)ins.component;
                    else if (ai >= 0) r.ids[0] = arc_base(ai) + 4 + (uint32_t)ins.component;
                    else return undefined(ins.l[0]);
                    r.p0 = ins.value;
                    push(r, kNaN);
                } break;
                case Ins::Vertical:
                case Ins::Horizontal:
                case Ins::PointsCoincident:
                case Ins::Distance: {
                    if (!point_for(ins.l[0], a)) return undefined(ins.l[0]);
                    if (!point_for(ins.l[1], b)) return undefined(ins.l[1]);
                    ezpz_constraint_t r = blank(ins.op == Ins::Vertical ? EZPZ_K_VERTICAL
                                                : ins.op == Ins::Horizontal ? EZPZ_K_HORIZONTAL
                                                : ins.op == Ins::PointsCoincident ? EZPZ_K_POINTS_COINCIDENT : EZPZ_K_DISTANCE);
                    set_pts(r, 0, a);
                    set_pts(r, 2, b);
                    if (ins.op == Ins::Distance) r.p0 = ins.value;
                    push(r, kNaN);
                } break;
                case Ins::PointArcCoincident: {  // point_arc_coincident(point, arc)
                    if (!point_for(ins.l[0], a)) return undefined(ins.l[0]);
                    ezpz_constraint_t r = blank(EZPZ_K_POINT_ARC_COINCIDENT);
                    if (!arc_of(ins.l[1], r, bad)) return undefined(bad);
                    set_pts(r, 6, a);
                    push(r, kNaN);
                } break;
                case Ins::Midpoint: {  // midpoint(p0, p1, mp)
                    if (!point_for(ins.l[0], a)) return undefined(ins.l[0]);
                    if (!point_for(ins.l[1], b)) return undefined(ins.l[1]);
                    if (!point_for(ins.l[2], c2)) return undefined(ins.l[2]);
                    ezpz_constraint_t r = blank(EZPZ_K_MIDPOINT);
                    set_pts(r, 0, a);
                    set_pts(r, 2, b);
                    set_pts(r, 4, c2);
                    push(r, kNaN);
                } break;
                case Ins::Symmetric: {  // symmetric(lineP, lineQ, a, b); a, b resolved first (executor.rs:347-359)
                    if (!point_for(ins.l[2], c2)) return undefined(ins.l[2]);
                    if (!point_for(ins.l[3], d2)) return undefined(ins.l[3]);
                    if (!point_for(ins.l[0], a)) return undefined(ins.l[0]);
                    if (!point_for(ins.l[1], b)) return undefined(ins.l[1]);
                    ezpz_constraint_t r = blank(EZPZ_K_SYMMETRIC);
                    set_pts(r, 0, a);
                    set_pts(r, 2, b);
                    set_pts(r, 4, c2);
                    set_pts(r, 6, d2);
                    push(r, kNaN);
                } break;
                case Ins::Parallel:
                case Ins::Perpendicular:
                case Ins::AngleLine:
                case Ins::LinesEqualLength: {
                    if (!point_for(ins.l[0], a)) return undefined(ins.l[0]);
                    if (!point_for(ins.l[1], b)) return undefined(ins.l[1]);
                    if (!point_for(ins.l[2], c2)) return undefined(ins.l[2]);
                    if (!point_for(ins.l[3], d2)) return undefined(ins.l[3]);
                    ezpz_constraint_t r = blank(ins.op == Ins::LinesEqualLength ? EZPZ_K_LINES_EQUAL_LENGTH : EZPZ_K_LINES_AT_ANGLE);
                    set_pts(r, 0, a);
                    set_pts(r, 2, b);
                    set_pts(r, 4, c2);
                    set_pts(r, 6, d2);
                    double adeg = kNaN;
                    if (ins.op == Ins::Parallel) {  // rotation_for_angle_kind (constraints.rs:2641-2647)
                        r.flags = EZPZ_ANGLE_PARALLEL;
                        r.p0 = 1.0;
                        r.p1 = 0.0;
                    } else if (ins.op == Ins::Perpendicular) {
                        r.flags = EZPZ_ANGLE_PERPENDICULAR;
                        r.p0 = 0.0;
                        r.p1 = 1.0;
                    } else if (ins.op == Ins::AngleLine) {
                        r.flags = EZPZ_ANGLE_OTHER;
                        const double kPi = 3.14159265358979323846;
                        const double rad = ins.degrees ? ins.value * (kPi / 180.0) : ins.value;  // f64::to_radians
                        double s, co;
                        ezm::ez_sincos(rad, s, co);
                        r.p0 = co;
                        r.p1 = s;
                        adeg = ins.degrees ? ins.value : ins.value * (180.0 / kPi);  // f64::to_degrees
                    }
                    push(r, adeg);
                } break;
            }
        }
        P->built = true;
    }
    if (cons) *cons = P->cons.data();
    if (n_cons) *n_cons = (uint32_t)P->cons.size();
    if (guesses) *guesses = P->guesses.data();
    if (n_vars) *n_vars = (uint32_t)P->guesses.size();
    return EZPZ_OK;
}

uint32_t ezpz_b200_problem_count(const ezpz_problem_t* p, int32_t kind) {
    if (!p) return 0;
    switch (kind) {
        case 0: return (uint32_t)p->points.size();
        case 1: return (uint32_t)p->circles.size();
        case 2: return (uint32_t)p->arcs.size();
        default: return 0;
    }
}

const char* ezpz_b200_problem_label(const ezpz_problem_t* p, int32_t kind, uint32_t index) {
    if (!p) return nullptr;
    const std::vector<std::string>* v = kind == 0 ? &p->points : kind == 1 ? &p->circles : kind == 2 ? &p->arcs : nullptr;
    if (!v || index >= v->size()) return nullptr;
    return (*v)[index].c_str();
}

int32_t ezpz_b200_problem_angles_deg(const ezpz_problem_t* p, const double** angles_deg) {
    if (!p || !angles_deg || !p->built) return EZPZ_ERR_INVALID_ARGUMENT;
    *angles_deg = p->angles_deg.data();
    return EZPZ_OK;
}

}  // extern "C"
