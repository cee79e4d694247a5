#!/usr/bin/env python
"""Dev-time generator for the discrete process-noise polynomial Q_d (test infrastructure + codegen).

The reference's Q_d (reference: src/x/ekf/propagator.cpp:207-840, Propagator::discreteProcessNoiseCov)
is a machine-generated symbolic polynomial, *not* a clean closed form of the Van-Loan integral
(e.g. several rows reuse the (0,0) entry of C[a]x where a consistent derivation would use other
entries, and the matrix is not symmetric).  Bit-level parity with the reference filter therefore
requires evaluating the *same polynomial*; it cannot be re-derived from the continuous-time model.

This tool (run only in the build container, where /root/reference is mounted) parses that function
into a sympy expression DAG, re-expresses it over the rotation matrix C = C(q) (instead of raw
quaternion products), the unbiased IMU sample (w, a), dt and the four noise densities, re-runs
common-subexpression elimination, and emits
  * oracle/qd_poly.py                    (numpy fp64 oracle evaluation)
  * x_multi_agent_b200/csrc/qd_poly.cuh  (device/host inline fp64 evaluation used by the kernels)
Both are checked against the reference function compiled in place (oracle/_ref/libxref_qd.so) by
tests/test_oracle_propagator.py.
"""
import re
import sys
from pathlib import Path

import sympy as sp

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/x/ekf/propagator.cpp")
ROOT = Path(__file__).resolve().parents[2]


def load_body():
    lines = REF.read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("CoreCovMatrix Propagator::discreteProcessNoiseCov"))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("}"))
    body = " ".join(l.strip() for l in lines[start:end])
    body = body[body.index("{") + 1:]
    return [s.strip() for s in body.split(";") if s.strip()]


def to_sympy_src(expr):
    expr = re.sub(r"\((-?\d+)\.0 / (\d+)\.0\)", r"Rational(\1,\2)", expr)
    expr = re.sub(r"(?<![\w.])(\d+)\.0(?![\w.])", r"\1", expr)
    return expr


def main():
    stmts = load_body()
    # Rotation-matrix entries, body->global, C = C(q) with q = (q1=w, q2=x, q3=y, q4=z).
    C = sp.Matrix(3, 3, lambda i, j: sp.Symbol(f"C{i}{j}"))
    w = sp.symbols("w0 w1 w2")
    a = sp.symbols("a0 a1 a2")
    dt, n_w, n_bw, n_a, n_ba = sp.symbols("dt n_w n_bw n_a n_ba")
    q1, q2, q3, q4 = sp.symbols("q1 q2 q3 q4")
    # exact quaternion polynomials of C (Eigen toRotationMatrix convention)
    Cq = {
        (0, 0): q1*q1 + q2*q2 - q3*q3 - q4*q4, (0, 1): 2*(q2*q3 - q1*q4), (0, 2): 2*(q2*q4 + q1*q3),
        (1, 0): 2*(q2*q3 + q1*q4), (1, 1): q1*q1 - q2*q2 + q3*q3 - q4*q4, (1, 2): 2*(q3*q4 - q1*q2),
        (2, 0): 2*(q2*q4 - q1*q3), (2, 1): 2*(q3*q4 + q1*q2), (2, 2): q1*q1 - q2*q2 - q3*q3 + q4*q4,
    }
    env = {"Rational": sp.Rational, "dt": dt, "n_w": n_w, "n_bw": n_bw, "n_a": n_a, "n_ba": n_ba,
           "q1": q1, "q2": q2, "q3": q3, "q4": q4,
           "e_w1": w[0], "e_w2": w[1], "e_w3": w[2], "e_a1": a[0], "e_a2": a[1], "e_a3": a[2]}
    idx = {"kIdxP": 0, "kIdxV": 3, "kIdxQ": 6, "kIdxBw": 9, "kIdxBa": 12}
    Q = sp.zeros(15, 15)

    def lift(e):
        """Replace a pure-quaternion quadratic by +-C_ij when it is one (keeps the DAG over C)."""
        ee = sp.expand(e)
        if ee.free_symbols and ee.free_symbols <= {q1, q2, q3, q4}:
            for (i, j), poly in Cq.items():
                if sp.expand(ee - poly) == 0:
                    return C[i, j]
                if sp.expand(ee + poly) == 0:
                    return -C[i, j]
        return e

    for s in stmts:
        m = re.match(r"const double (.*)", s)
        if m:
            for part in m.group(1).split(","):
                if "=" not in part:
                    continue
                name, expr = part.split("=", 1)
                name = name.strip()
                if re.match(r"q\d$|e_[wa]\d$", name):
                    continue  # q1 = q.w() ... handled by env
                env[name] = lift(eval(to_sympy_src(expr.strip()), {}, env))
            continue
        m = re.match(r"q_d\((\w+) \+ (\d), (\w+) \+ (\d)\) = (.*)", s)
        if m:
            r = idx[m.group(1)] + int(m.group(2))
            c = idx[m.group(3)] + int(m.group(4))
            val = eval(to_sympy_src(m.group(5)), {}, env)
            # one entry spells a C entry inline from quaternion products
            val = val.subs(sp.expand(2*q1*q2 - 2*q3*q4), -C[1, 2]) if val.has(q1) else val
            if val.has(q1) or val.has(q2) or val.has(q3) or val.has(q4):
                val = sp.factor_terms(val)
                val = val.xreplace({q1*q2*2 - q3*q4*2: -C[1, 2]})
            assert not (val.free_symbols & {q1, q2, q3, q4}), (r, c, val)
            Q[r, c] = val
            continue
        if s.startswith("CoreCovMatrix q_d") or s.startswith("return"):
            continue
        raise SystemExit(f"unparsed statement: {s[:80]}")

    nz = [(r, c) for r in range(15) for c in range(15) if Q[r, c] != 0]
    repl, red = sp.cse([Q[r, c] for r, c in nz], symbols=sp.numbered_symbols("s"), optimizations="basic")
    print(f"nonzero entries: {len(nz)}, cse temporaries: {len(repl)}", file=sys.stderr)

    hdr = ("Q_d polynomial of the reference filter (reference: src/x/ekf/propagator.cpp:207-840), re-expressed over "
           "C=C(q), w, a, dt and the noise densities and re-factored by oracle/tools/gen_qd.py. GENERATED - do not edit.")
    # ---- python
    py = [f'"""{hdr}\nTest infrastructure only (oracle)."""', "import numpy as np", "", "",
          "def qd_poly(dt, C, w, a, n_w, n_bw, n_a, n_ba):",
          '    """C: 3x3 rotation (Eigen toRotationMatrix of q1), w/a: unbiased gyro/accel of the newer state."""',
          "    (C00, C01, C02), (C10, C11, C12), (C20, C21, C22) = C",
          "    w0, w1, w2 = w", "    a0, a1, a2 = a"]
    for sym, e in repl:
        py.append(f"    {sym} = {sp.pycode(e)}")
    py.append("    Q = np.zeros((15, 15))")
    for (r, c), e in zip(nz, red):
        py.append(f"    Q[{r}, {c}] = {sp.pycode(e)}")
    py.append("    return Q")
    (ROOT / "oracle" / "qd_poly.py").write_text("\n".join(py) + "\n")
    # ---- CUDA / C++
    cu = [f"// {hdr}", "#pragma once", "#ifndef XB_HD", "#ifdef __CUDACC__", "#define XB_HD __host__ __device__ __forceinline__",
          "#else", "#define XB_HD inline", "#endif", "#endif",
          "// C row-major 3x3; Q row-major 15x15 (must be zero-initialised by the caller).",
          "XB_HD void xb_qd_poly(double dt, const double* C, const double* w, const double* a, double n_w, double n_bw,",
          "                      double n_a, double n_ba, double* Q) {",
          "  const double C00 = C[0], C01 = C[1], C02 = C[2], C10 = C[3], C11 = C[4], C12 = C[5], C20 = C[6], C21 = C[7], C22 = C[8];",
          "  const double w0 = w[0], w1 = w[1], w2 = w[2], a0 = a[0], a1 = a[1], a2 = a[2];"]
    for sym, e in repl:
        cu.append(f"  const double {sym} = {sp.ccode(e)};")
    for (r, c), e in zip(nz, red):
        cu.append(f"  Q[{r * 15 + c}] = {sp.ccode(e)};")
    cu.append("}")
    out = ROOT / "x_multi_agent_b200" / "csrc"
    out.mkdir(parents=True, exist_ok=True)
    (out / "qd_poly.cuh").write_text("\n".join(cu) + "\n")

    pos = {sym: i for i, (sym, _) in enumerate(repl)}
    ops = [sp.count_ops(e) + 1 for _, e in repl]
    ecost = [sp.count_ops(e) + 1 for e in red]
    # ---- level-scheduled form: every statement of the global CSE exactly once (bit-identical to xb_qd_poly), temporaries in
    #      shared memory S[], statements of one dependency level spread over NWARP warps (one lane each); the caller puts a
    #      block barrier between levels.  Per warp this is ~1/5 of the instructions of a partition above: the kernels run
    #      this code once per launch on cold instruction caches, where the instruction count IS the latency.  (An earlier
    #      form split the ENTRIES into 7 groups with a CSE per group: faster still, but a CSE per group re-associates
    #      cancelling terms -- 1e-10 relative differences in the small entries, amplified to 1e-7 by one ill-conditioned
    #      update of the 65-update parity sequence; with the global CSE per group every group recomputes ~95 of the 224
    #      temporaries.)
    NWARP = 15
    lvl_t = []
    for i, (sym, e) in enumerate(repl):
        d = [pos[fs] for fs in e.free_symbols if fs in pos]
        lvl_t.append(1 + max([lvl_t[j] for j in d], default=-1))
    lvl_e = [1 + max([lvl_t[pos[fs]] for fs in e.free_symbols if fs in pos], default=-1) for e in red]
    nlev = max(max(lvl_t), max(lvl_e)) + 1
    sched = [[[] for _ in range(NWARP)] for _ in range(nlev)]
    loadw = [[0] * NWARP for _ in range(nlev)]

    def to_s(code):
        return re.sub(r"\bs(\d+)\b", r"S[\1]", code)

    def place(level, cost, text):
        w = loadw[level].index(min(loadw[level]))
        sched[level][w].append(text)
        loadw[level][w] += cost

    for i in sorted(range(len(repl)), key=lambda i: -ops[i]):
        place(lvl_t[i], ops[i], f"S[{i}] = {to_s(sp.ccode(repl[i][1]))};")
    # entries have no dependants: an entry goes to the least loaded level at or after its own
    for i in sorted(range(len(nz)), key=lambda i: -ecost[i]):
        r, c = nz[i]
        best = min(range(lvl_e[i], nlev), key=lambda L: min(loadw[L]))
        place(best, ecost[i], f"Q[{r * 15 + c}] = {to_s(sp.ccode(red[i]))};")
    assert all(str(sym) == f"s{i}" for i, (sym, _) in enumerate(repl))
    cl = [f"// {hdr}", "// Level-scheduled form: xb_qd_poly_warp(w, ...) is executed by ALL lanes of warp w (0..XB_QD_NWARP-1) of a CTA, redundantly",
          "// (the polynomial is scalar); it runs that warp's statements of every dependency level in turn, with a named barrier",
          "// over the XB_QD_NWARP warps between levels (bar.sync 1).  Temporaries live in S[XB_QD_NTEMP] (shared memory).  Every",
          "// statement of xb_qd_poly (qd_poly.cuh) appears exactly once, with the same text: bit-identical results.  A warp's",
          "// statements are contiguous in the code (sequential instruction fetch: the kernels run this once per launch).",
          "#pragma once", f"#define XB_QD_NLEVEL {nlev}", f"#define XB_QD_NWARP {NWARP}", f"#define XB_QD_NTEMP {len(repl)}",
          f'#define XB_QD_BAR() asm volatile("bar.sync 1, {32 * NWARP};" ::: "memory")',
          "__device__ __noinline__ void xb_qd_poly_warp(int w, double dt, const double* C, const double* wv, const double* a,",
          "                                             double n_w, double n_bw, double n_a, double n_ba, double* S, double* Q) {",
          "  const double C00 = C[0], C01 = C[1], C02 = C[2], C10 = C[3], C11 = C[4], C12 = C[5], C20 = C[6], C21 = C[7], C22 = C[8];",
          "  const double w0 = wv[0], w1 = wv[1], w2 = wv[2], a0 = a[0], a1 = a[1], a2 = a[2];",
          "  switch (w) {"]
    for w in range(NWARP):
        cl.append(f"    case {w}: {{")
        for L in range(nlev):
            cl += ["      " + t for t in sched[L][w]]
            if L + 1 < nlev:
                cl.append("      XB_QD_BAR();")
        cl += ["      break;", "    }"]
    for L in range(nlev):
        print(f"level {L}: {sum(len(x) for x in sched[L])} statements, max warp cost {max(loadw[L])}", file=sys.stderr)
    cl += ["    default: break;", "  }", "}"]
    (out / "qd_poly_levels.cuh").write_text("\n".join(cl) + "\n")


if __name__ == "__main__":
    main()
