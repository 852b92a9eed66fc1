#!/usr/bin/env python3
"""Generate the device code of the four ERK methods (flint_b200/csrc/erk_methods_gen.cuh).

Reads the Butcher tableaus through tools/gen_tableaus.py (i.e. from the reference's
src/ButcherTableaus.f90) and emits, per method, a struct with

    step<SYS,FMA,DENSE>()   one step attempt: all stages, candidate solution, error norm,
                            the evaluation the reference performs on acceptance and (DENSE)
                            the extra dense-output stages
    dense_coeffs<SYS,FMA>() dense-output coefficients (power basis) from the stage vectors

Design (B200-first, measured -- profiles/r01_notes.md):
  * one trajectory per thread, every stage vector in REGISTERS: compile-time stage
    indices, coefficients as literals (constant bank / immediates, warp-uniform);
  * the kernel is instruction-fetch bound if the step is fully unrolled (a DOP853 step
    with 12 inlined right-hand sides is >80 KB of SASS against a 32 KB L1.5 I-cache), so
    the stages run in ONE runtime loop whose body is `switch(stage){sums}` + a single
    inlined RHS + `switch(stage){store}`.  All lanes of a warp are always at the same
    stage, so the switches never diverge;
  * stage vectors are assigned to register SLOTS by liveness (a stage that is never
    read again gives its slot to a later one), e.g. DOP853 needs 12 slots for 16 stages;
  * structural zeros of the tableau are skipped at generation time.  This is exact:
    a partial sum that starts from +0.0 is never -0.0, so adding k*0 = +-0 never
    changes it (finite k).  Where the reference's loops start from 0.0 (generic
    stepint / IntpCoeff, src/ERKStepInt.f90:120-323) the generated code keeps the
    leading `0.0 +`; where its hand-unrolled code does not (DOP853, DOP54) neither
    does this.  DOP54's stage-7 row keeps the explicit `k2*0` term the reference
    has (src/ERKStepInt.f90:615-616);
  * operation order inside every sum is the reference's (ascending stage index),
    so strict mode is bit-identical to the oracle and FMA mode to the FMA oracle.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_tableaus as GT  # noqa: E402


def hx(v):
    return float(v).hex()


def aidx(i, j):
    """packed index (0-based) of a_ij, 1-based stage i, column j (src/ERKStepInt.f90:162)"""
    return (i - 1) * (i - 2) // 2 + j - 1


class ConstPool:
    """Coefficients travel in the KERNEL PARAMETER block (KArgs::coef, constant bank 0), so every
    DFMA/DMUL takes its coefficient as a c[0x0][off] operand.  Two alternatives were measured and
    dropped: literals (nvcc materialises each 64-bit immediate with two UMOVs, ~25 % more instructions)
    and `__constant__` arrays (the optimiser hoists those loads out of the lane loop as loop
    invariants and ptxas then SPILLS the coefficients to local memory: 12 % of all stall samples sat
    on the reloads, profiles/r01_notes.md).  The host copies kCH_<method> into KArgs::coef at launch."""

    def __init__(self, name):
        self.name, self.vals, self.index = name, [], {}

    def ref(self, v):
        key = float(v).hex()
        if key not in self.index:
            self.index[key] = len(self.vals)
            self.vals.append(float(v))
        return "C[%d]" % self.index[key]

    def decl(self):
        body = ",\n    ".join(", ".join(hx(v) for v in self.vals[i:i + 4]) for i in range(0, len(self.vals), 4))
        return "static const double %s[%d] = {\n    %s\n};\n" % (self.name, len(self.vals), body)


POOL = None


def cx(v):
    """coefficient operand: constant-bank reference"""
    return POOL.ref(v)


class Emit:
    def __init__(self):
        self.lines = []
        self.ind = 0

    def __call__(self, s=""):
        self.lines.append(("    " * self.ind + s) if s else "")

    def text(self):
        return "\n".join(self.lines) + "\n"


INF = 10 ** 9
# True : straight-line stages, the right-hand side is ONE __noinline__ function called per stage
# False: runtime stage loop with switch(stage) around a single inlined RHS
UNROLLED = True


class Plan:
    """Iteration plan of one method: which stage is evaluated in which loop iteration, which
    stage vectors each iteration reads, and the register slot of every stage vector."""

    def __init__(self, S, M):
        p = M.lower()
        self.M = M
        self.s, self.sint = S[p + "_s"], S[p + "_sint"]
        self.a, self.b, self.c, self.e = S[p + "_a"], S[p + "_b"], S[p + "_c"], S[p + "_e"]
        self.d, self.dinz = S[p + "_d"], S[p + "_di_nz"]
        self.fsal = bool(S[p + "_fsal"])
        self.pstar = S[p + "_pstar"]
        self.generic = M.startswith("Verner")
        self.bhh = S.get("dop853_bhh")
        s, sint = self.s, self.sint
        # iterations: (kind, stage)
        #   'stage': evaluate k_stage at the stage's own argument
        #   'acc'  : evaluate F(X0+h, Ynew) (DOP853: this IS k13; Verner98R: pseudo stage 0 = next F0)
        its = []
        if M == "DOP853":
            its += [("stage", i) for i in range(2, 13)] + [("acc", 13)] + [("stage", i) for i in (14, 15, 16)]
            self.n_int = 12
        elif M == "DOP54":
            its += [("stage", i) for i in range(2, 8)]
            self.n_int = 6
        elif self.fsal:
            its += [("stage", i) for i in range(2, s + 1)]
            self.n_int = sint - 1
        else:
            its += [("stage", i) for i in range(2, sint + 1)] + [("acc", 0)] + [("stage", i) for i in range(sint + 1, s + 1)]
            self.n_int = sint
        self.its = its
        # ---- uses of every stage vector ----
        uses = {j: [] for j in range(0, s + 1)}
        for it, (kind, st) in enumerate(its):
            row = st if kind == "stage" else (13 if M == "DOP853" else None)
            if row:
                for j in range(1, row):
                    if self.a[aidx(row, j)] != 0.0 or (M == "DOP54" and row == 7):
                        uses[j].append(it)
            if kind == "acc" and M == "Verner98R":
                for j in range(1, sint + 1):
                    if self.b[j - 1] != 0.0 or self.e[j - 1] != 0.0:
                        uses[j].append(it)
            if kind == "acc" and M == "DOP853":
                for j in range(1, sint + 1):
                    if self.e[j - 1] != 0.0:
                        uses[j].append(it)
            # FSAL generic with DENSE: the error norm is formed in front of the first dense stage
            if self.generic and self.fsal and kind == "stage" and st == sint + 1:
                for j in range(1, sint + 1):
                    if self.e[j - 1] != 0.0:
                        uses[j].append(it)
        # after the loop: error (DOP54 / FSAL generic), next F0, stiffness pair, dense coefficients
        post = set()
        if M == "DOP54" or (self.generic and self.fsal):
            post |= {j for j in range(1, sint + 1) if self.e[j - 1] != 0.0}
        post |= {sint, sint - 1}                     # CheckStiffness: k(:,sint) - k(:,sint-1)
        post |= set(self.dinz)                       # dense coefficients
        post.add(1)
        if M == "DOP853":
            post.add(13)
        if M == "DOP54":
            post.add(7)
        if M == "Verner98R":
            post.add(0)
        for j in post:
            uses[j].append(INF)
        self.last = {j: (max(u) if u else -1) for j, u in uses.items()}
        # ---- slot allocation (stage 1 = F0 copy, slot 0, live for ever) ----
        slot = {1: 0}
        free, nslot = [], 1
        busy_until = {0: INF}
        for it, (kind, st) in enumerate(its):
            for sl, until in list(busy_until.items()):
                if until < it:      # owner is not read at or after this iteration
                    free.append(sl)
                    del busy_until[sl]
            free.sort()
            sl = free.pop(0) if free else nslot
            if sl == nslot:
                nslot += 1
            slot[st] = sl
            busy_until[sl] = max(self.last[st], it)
        self.slot, self.nslot = slot, nslot

    def k(self, j):
        return "ks.template get<%d>(c)" % self.slot[j]


def emit_sum(E, terms, lead_zero, keep_zero_terms=False, var="acc"):
    """var = sum of k_j[c] * coef_j, left to right (from 0.0 when lead_zero).  For a component whose
    derivative is structurally +0.0 (SYS::fsrc(c) == -1) every product is +-0, so the sum is a zero whose
    sign is known here: -0.0 iff the sum does not start from 0.0 and every coefficient is negative
    ((-0) + (-0) = -0, anything else rounds to +0).  The branch folds after unrolling."""
    used = [(k, c) for k, c in terms if c != 0.0 or keep_zero_terms]
    if not used:
        E("double %s = 0.0;" % var)
        return
    zsign = "-0.0" if (not lead_zero and all(c < 0.0 for _, c in used)) else "0.0"
    E("double %s;" % var)
    E("if (SYS::fsrc(c) == -1) %s = %s;" % (var, zsign))
    E("else {")
    E.ind += 1
    first = True
    for kexpr, coef in used:
        if first:
            if lead_zero:
                E("%s = madd<FMA>(%s, %s, 0.0);" % (var, kexpr, cx(coef)))
            else:
                E("%s = __dmul_rn(%s, %s);" % (var, kexpr, cx(coef)))
            first = False
        else:
            E("%s = madd<FMA>(%s, %s, %s);" % (var, kexpr, cx(coef), var))
    E.ind -= 1
    E("}")


def stage_sums(E, P, row, dest, lead_zero, keep_zero_terms=False, save_sum=None, hname="h"):
    E("#pragma unroll")
    E("for (int c = 0; c < N; ++c) {")
    E.ind += 1
    # a component whose state AND derivative are identically +0.0 (SYS::zc) stays +0.0: (+0) + h*(+-0) = +0
    E("if (SYS::zc(c)) { %s[c] = 0.0; %scontinue; }" % (dest, ("%s[c] = 0.0; " % save_sum) if save_sum else ""))
    emit_sum(E, [(P.k(j), P.a[aidx(row, j)]) for j in range(1, row)], lead_zero, keep_zero_terms)
    if save_sum:
        E("%s[c] = acc;" % save_sum)
    E("%s[c] = madd<FMA>(%s, acc, Y0[c]);" % (dest, hname))
    E.ind -= 1
    E("}")


def emit_error(E, P):
    """error norm into Err (estimate == true path).

    The quotients by Sc are formed without branches: one refined reciprocal per component
    (arith.cuh: rcp_newton, div_rcp_nochk) and a range test that needs only Sc and the size of the
    quotient, because only q*q is used: a quotient below 2^-600 squares to +0 whatever its last bits or
    sign.  A failed test redoes all quotients with the plain operator out of line.  Components whose state and
    derivative are identically +0 (SYS::zc) contribute (+-0/Sc)^2 = +0 and are skipped (the host selects such a
    kernel only when ATol > 0 there, so Sc > 0); components with a structurally zero derivative (fsrc == -1)
    contribute +0 when Sc is positive and finite, else what 0/Sc gives."""
    M = P.M
    E("static_assert(!SYS::zc(0), \"component 0 starts the error sums\");")
    E("bool qok = true;")
    E("VecN<N> scv;")
    if M == "DOP853":
        E("VecN<N> t5, t3, q5, q3;")
    else:
        E("VecN<N> tn, qn;")
    E("#pragma unroll")
    E("for (int c = 0; c < N; ++c) {")
    E.ind += 1
    if M == "DOP853":
        E("if (SYS::zc(c)) { scv.v[c] = 1.0; t5.v[c] = t3.v[c] = q5.v[c] = q3.v[c] = 0.0; continue; }")
    else:
        E("if (SYS::zc(c)) { scv.v[c] = 1.0; tn.v[c] = qn.v[c] = 0.0; continue; }")
    E("const double sc = madd<FMA>(tol.rt(c), dmax(fabs(Y0[c]), fabs(Ynew[c])), tol.at(c));")
    E("scv.v[c] = sc;")
    if M == "DOP853":
        emit_sum(E, [(P.k(j + 1), P.e[j]) for j in range(P.sint)], lead_zero=False, var="t")
        E("t5.v[c] = t;")
        E("double u = madd<FMA>(%s, %s, bsum[c]);" % (cx(-P.bhh[0]), P.k(1)))
        E("u = madd<FMA>(%s, %s, u);" % (cx(-P.bhh[1]), P.k(9)))
        E("u = madd<FMA>(%s, %s, u);" % (cx(-P.bhh[2]), P.k(12)))
        E("t3.v[c] = u;")
        E("if (SYS::fsrc(c) == -1) { q5.v[c] = q3.v[c] = zero_over(sc); continue; }")
        E("const double isc = rcp_newton(sc);")
        E("q5.v[c] = div_rcp_nochk(t, sc, isc); q3.v[c] = div_rcp_nochk(u, sc, isc);")
        E("qok = qok & moderate(sc) & below_2p600(q5.v[c]) & below_2p600(q3.v[c]);")
    else:
        emit_sum(E, [(P.k(j + 1), P.e[j]) for j in range(P.sint)], lead_zero=P.generic, var="t")
        E("tn.v[c] = h * t;")
        E("if (SYS::fsrc(c) == -1) { qn.v[c] = zero_over(sc); continue; }")
        E("qn.v[c] = div_rcp_nochk(tn.v[c], sc, rcp_newton(sc));")
        E("qok = qok & moderate(sc) & below_2p600(qn.v[c]);")
    E.ind -= 1
    E("}")
    if M == "DOP853":
        E("if (!qok) { q5 = quotients_plain<N>(t5, scv); q3 = quotients_plain<N>(t3, scv); }")
        E("double Es = 0.0, E3 = 0.0;")
        E("#pragma unroll")
        E("for (int c = 0; c < N; ++c) if (!SYS::zc(c)) Es = (c == 0) ? __dmul_rn(q5.v[c], q5.v[c]) : madd<FMA>(q5.v[c], q5.v[c], Es);")
        E("#pragma unroll")
        E("for (int c = 0; c < N; ++c) if (!SYS::zc(c)) E3 = (c == 0) ? __dmul_rn(q3.v[c], q3.v[c]) : madd<FMA>(q3.v[c], q3.v[c], E3);")
        E("double den = madd<FMA>(0.01, E3, Es);")
        E("if (den == 0.0) den = 1.0;")
        E("Err = (fabs(h) * Es) * sqrt(1.0 / ((double)N * den));")
    else:
        E("if (!qok) qn = quotients_plain<N>(tn, scv);")
        E("double sm = 0.0;")
        E("#pragma unroll")
        E("for (int c = 0; c < N; ++c) if (!SYS::zc(c)) sm = (c == 0) ? __dmul_rn(qn.v[c], qn.v[c]) : madd<FMA>(qn.v[c], qn.v[c], sm);")
        E("Err = sqrt(sm / (double)N);")


def emit_iteration_pre(E, P, it, kind, st):
    M, sint, fsal = P.M, P.sint, P.fsal
    if kind == "stage":
        hn = "hd" if st > sint else "h"
        if M == "DOP853":
            dest = "Ypre" if st == 12 else "yt"
            stage_sums(E, P, st, dest, lead_zero=False, hname=hn)
        elif M == "DOP54":
            dest = "Ypre" if st == 6 else ("Ynew" if st == 7 else "yt")
            stage_sums(E, P, st, dest, lead_zero=False, keep_zero_terms=True)
        else:
            if st <= sint:
                dest = "Ypre" if st == sint - 1 else ("Ynew" if st == sint else "yt")
            else:
                dest = "yt"
            if P.generic and fsal and st == sint + 1:
                E("if (estimate) {   /* FSAL generic + DENSE: the error norm precedes the first dense stage */")
                E.ind += 1
                emit_error(E, P)
                E.ind -= 1
                E("}")
            stage_sums(E, P, st, dest, lead_zero=True, hname=hn)
        if dest != "yt":
            E("#pragma unroll")
            E("for (int c = 0; c < N; ++c) yt[c] = %s[c];" % dest)
        if M == "DOP54" and st == 7:
            E("xs = X0 + h;")
        else:
            E("xs = madd<FMA>(%s, %s, X0);" % (hn, cx(P.c[st - 2])))
    else:
        if M == "DOP853":
            E("/* row 13 (= b): candidate solution */")
            stage_sums(E, P, 13, "Ynew", lead_zero=False, save_sum="bsum")
        else:
            E("/* non-FSAL: Ynew = Y0 + h*matmul(k, b)  (src/ERKStepInt.f90:198) */")
            E("#pragma unroll")
            E("for (int c = 0; c < N; ++c) {")
            E.ind += 1
            E("if (SYS::zc(c)) { Ynew[c] = 0.0; continue; }")
            emit_sum(E, [(P.k(j + 1), P.b[j]) for j in range(sint)], lead_zero=True)
            E("Ynew[c] = madd<FMA>(h, acc, Y0[c]);")
            E.ind -= 1
            E("}")
        E("if (estimate) {")
        E.ind += 1
        emit_error(E, P)
        E.ind -= 1
        E("}")
        E("#pragma unroll")
        E("for (int c = 0; c < N; ++c) yt[c] = Ynew[c];")
        E("xs = X0 + h;")


def gen_method(E_out, S, M):
    global POOL
    POOL = ConstPool("kCH_%s" % M)
    E = Emit()
    P = Plan(S, M)
    p = M.lower()
    U = M.upper()
    s, sint, fsal, pstar = P.s, P.sint, P.fsal, P.pstar
    if M in ("DOP853", "DOP54") or fsal:
        fc_acc, fc_rej = sint - 1, sint - 2
    else:
        fc_acc, fc_rej = sint, sint - 1
    fc_dense = 0 if M == "DOP54" else s - sint
    n_all = len(P.its)
    err_in_loop = M in ("DOP853", "Verner98R")      # the error norm is formed in the `acc` iteration

    E("/* ================= %s : src/ButcherTableaus.f90; stepping as %s ================= */" % (
        M, {"DOP853": "src/ERKStepInt.f90:327-563", "DOP54": "src/ERKStepInt.f90:568-700"}.get(M, "src/ERKStepInt.f90:120-323")))
    E("struct Method%s {" % M)
    E.ind += 1
    E("static constexpr int ID = ERK_%s;" % U)
    E("static constexpr int S = %d, SINT = %d, P = %d, Q = %d, PSTAR = %d;" % (s, sint, S[p + "_p"], S[p + "_q"], pstar))
    E("static constexpr bool FSAL = %s;" % ("true" if fsal else "false"))
    E("static constexpr int FC_ACC = %d, FC_REJ = %d, FC_DENSE = %d;   /* FCalls accounting, SURVEY.md App. A.1 */" % (fc_acc, fc_rej, fc_dense))
    E("static constexpr bool YSINT_ASSIGNED = %s;   /* quirk Q3: generic FSAL path never sets Ysint */" % (
        "false" if (P.generic and fsal) else "true"))
    E("/* stage vectors live in NSLOT register slots (liveness-based reuse): stage -> slot")
    E(" *   " + ", ".join("k%d->%d" % (j, P.slot[j]) for j in sorted(P.slot) if j > 0) + (", F0next->%d" % P.slot[0] if 0 in P.slot else "") + " */")
    E("static constexpr int NSLOT = %d;" % P.nslot)
    E("/* slots the hot kernel keeps in shared memory instead of registers (erk_kernel.cuh: KStore) */")
    E("static constexpr unsigned SMEM_SLOTS = FLINT_SMEM_SLOTS_%s;" % U)
    E("static constexpr int NIT = %d, NIT_DENSE = %d;   /* loop iterations: integration / with dense stages */" % (P.n_int, n_all))
    E("static constexpr int SLOT_KSINT = %d, SLOT_KSINT_M1 = %d;   /* CheckStiffness operands */" % (P.slot[sint], P.slot[sint - 1]))
    f0new_stage = {"DOP853": 13, "DOP54": 7, "Verner65E": sint, "Verner98R": 0}[M]
    E("static constexpr int SLOT_F0NEW = %d;" % P.slot[f0new_stage])
    E("static constexpr int NCOEF = NCOEF_%s;" % M)
    E("static void load_coefficients(double (&dst)[FLINT_NCOEF]) { for (int i = 0; i < NCOEF; ++i) dst[i] = kCH_%s[i]; }" % M)
    E()
    # ------------------------------------------------------------------ step
    E("/* One step attempt (and, DENSE, the extra dense-output stages).  F0 = derivative at (X0,Y0).")
    E(" * h drives the integration stages, hd the dense stages (they differ only when an event action")
    E(" * truncated the step: reference quirk Q6 rebuilds the dense stages with the new h and the old k's).")
    E(" * Out: ks (stage vectors), Ynew (candidate), Ypre (state of stage sint-1, for CheckStiffness),")
    E(" * Err (0 when !estimate).  The evaluation at (X0+h, Ynew) is always made (it is the next F0")
    E(" * when the step is accepted and costs 1/%d of the attempt when it is not).  HOT: the call comes from the lane loop" % (P.n_int))
    E(" * of erk_step_kernel (right-hand sides may be inlined and may defer their range check, arith.cuh: rhs_eval). */")
    E("template <class SYS, bool FMA, bool DENSE, bool HOT, class KS>")
    E("static __device__ __forceinline__ void step(const SYS& sys, double X0, const double (&Y0)[SYS::N], const double (&F0)[SYS::N],")
    E("        double h, double hd, KS& ks, double (&Ynew)[SYS::N], double (&Ypre)[SYS::N], const Tol& tol,")
    E("        bool estimate, double& Err, const double (&C)[FLINT_NCOEF])")
    E("{")
    E.ind += 1
    E("constexpr int N = SYS::N;")
    E("double yt[N], xs = X0;")
    if not UNROLLED:
        E("double f[N];")
    if M == "DOP853":
        E("double bsum[N];")
    E("#pragma unroll")
    E("for (int c = 0; c < N; ++c) yt[c] = Y0[c];")
    E("ks.template put<0>(F0);")
    E("Err = 0.0;")
    if UNROLLED:
        for it, (kind, st) in enumerate(P.its):
            if it == P.n_int:
                E("if constexpr (DENSE) {")
                E.ind += 1
            E("{   /* %s */" % (("stage %d" % st) if kind == "stage" else "evaluation at (X0+h, Ynew)"))
            E.ind += 1
            emit_iteration_pre(E, P, it, kind, st)
            E("{ double fv[N]; rhs_eval<SYS, FMA, FLINT_RHS_COPY(%d, %d), HOT>(sys, yt, fv); ks.template put<%d>(fv); }" % (it, len(P.its), P.slot[st]))
            E.ind -= 1
            E("}")
        if len(P.its) > P.n_int:
            E.ind -= 1
            E("}")
    else:
        E("#pragma unroll 1")
        E("for (int it = 0; it < nit; ++it) {")
        E.ind += 1
        E("switch (it) {")
        for it, (kind, st) in enumerate(P.its):
            E("case %d: {   /* %s */" % (it, ("stage %d" % st) if kind == "stage" else "evaluation at (X0+h, Ynew)"))
            E.ind += 1
            emit_iteration_pre(E, P, it, kind, st)
            E.ind -= 1
            E("} break;")
        E("default: break;")
        E("}")
        E("sys.template rhs<FMA>(xs, yt, f);")
        E("switch (it) {")
        for it, (kind, st) in enumerate(P.its):
            E("case %d:" % it)
            E("#pragma unroll")
            E("    ks.template put<%d>(f);" % P.slot[st])
            E("    break;")
        E("default: break;")
        E("}")
        E.ind -= 1
        E("}")
    if not err_in_loop:
        if P.generic and fsal:
            E("if (!DENSE && estimate) {   /* with DENSE the norm was formed before the first dense stage */")
        else:
            E("if (estimate) {")
        E.ind += 1
        emit_error(E, P)
        E.ind -= 1
        E("}")
    E.ind -= 1
    E("}")
    E()
    # ------------------------------------------------------------------ dense coefficients
    d, dinz = P.d, P.dinz
    rows = d["shape"][0]
    E("/* Dense-output coefficients B[j][c] of theta^j, j = 0..PSTAR (all n states), from the stage vectors. */")
    E("template <class SYS, bool FMA, class KS>")
    E("static __device__ __forceinline__ void dense_coeffs(const KS& ks, const double (&Y0)[SYS::N], double h,")
    E("        const double (&Y1)[SYS::N], double (&B)[PSTAR + 1][SYS::N], const double (&C)[FLINT_NCOEF])")
    E("{")
    E.ind += 1
    E("constexpr int N = SYS::N;")
    E("#pragma unroll")
    E("for (int c = 0; c < N; ++c) {")
    E.ind += 1
    if M in ("DOP853", "DOP54"):
        last = 13 if M == "DOP853" else 7
        E("const double cont1 = Y1[c] - Y0[c];")
        E("const double cont2 = madd<FMA>(h, %s, -cont1);" % P.k(1))
        E("const double cont3 = madd<FMA>(-h, %s, cont1) - cont2;" % P.k(last))
        ncol = 4 if M == "DOP853" else 1
        for col in range(ncol):
            emit_sum(E, [(P.k(dinz[r]), d["data"][r + rows * col]) for r in range(rows)], lead_zero=False, var="a%d" % (col + 4))
            E("const double cont%d = h * a%d;" % (col + 4, col + 4))
        E("B[0][c] = Y0[c];")
        E("B[1][c] = cont1 + cont2;")
        E("B[2][c] = (cont3 + cont4) - cont2;")
        if M == "DOP853":
            E("B[3][c] = ((cont5 + cont6) - 2.0 * cont4) - cont3;")
            E("B[4][c] = (madd<FMA>(-3.0, cont6, cont7) - 2.0 * cont5) + cont4;")
            E("B[5][c] = madd<FMA>(3.0, cont6, -3.0 * cont7) + cont5;")
            E("B[6][c] = madd<FMA>(3.0, cont7, -cont6);")
            E("B[7][c] = -cont7;")
        else:
            E("B[3][c] = -2.0 * cont4 - cont3;")
            E("B[4][c] = cont4;")
    else:
        E("(void)Y1;")
        E("B[0][c] = Y0[c];")
        for j in range(1, pstar + 1):
            E("{")
            E.ind += 1
            emit_sum(E, [(P.k(dinz[r]), d["data"][r + rows * (j - 1)]) for r in range(rows)], lead_zero=True, var="cont")
            E("B[%d][c] = h * cont;" % j)
            E.ind -= 1
            E("}")
    E.ind -= 1
    E("}")
    E.ind -= 1
    E("}")
    E.ind -= 1
    E("};")
    E()
    E_out.lines.append("/* coefficients of %s (deduplicated values; copied into KArgs::coef at launch) */" % M)
    E_out.lines.extend(POOL.decl().rstrip("\n").split("\n"))
    E_out.lines.append("constexpr int NCOEF_%s = %d;" % (M, len(POOL.vals)))
    E_out.lines.append("static_assert(NCOEF_%s <= FLINT_NCOEF, \"KArgs::coef too small\");" % M)
    E_out.lines.extend(E.lines)


HEAD = """/* GENERATED by tools/gen_kernels.py -- DO NOT EDIT.
 * Device code of the four ERK methods of the reference (coefficients from
 * src/ButcherTableaus.f90; hex floats are the exact real64 values).  Coefficients are
 * literals: nvcc places them in the constant bank (c[0x2]) or as immediates, i.e. "Butcher
 * coefficients in constant memory" with warp-uniform broadcast and no explicit loads.
 * See tools/gen_kernels.py for the structure (stage loop + switch, register slots).
 */
#ifndef FLINT_B200_ERK_METHODS_GEN_CUH
#define FLINT_B200_ERK_METHODS_GEN_CUH

#include "arith.cuh"
#include "kargs.h"

namespace flint {

"""

TAIL = """}  // namespace flint
#endif
"""


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    S = GT.load_symbols(ref)
    E = Emit()
    for M in ("DOP853", "DOP54", "Verner98R", "Verner65E"):
        gen_method(E, S, M)
    path = os.path.join(root, "flint_b200", "csrc", "erk_methods_gen.cuh")
    with open(path, "w") as f:
        f.write(HEAD + E.text() + TAIL)
    print("wrote", path, len(E.lines), "lines")


if __name__ == "__main__":
    main()
