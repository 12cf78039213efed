// climt_b200 -- Emanuel moist convection (CONVECT 4.3c): warp-cooperative device code (also compiled for the host by tests/emul).
//
// Replaces the reference's per-column Fortran routine and its callers
//   SUBROUTINE CONVECT / TLIFT      climt/_lib/emanuel/convect43c.f90:146-1219
//   column loop                     climt/_components/emanuel/_emanuel_convection.pyx:96-201
//   bolton_q_sat                    climt/_core/util.py:177-180          (saturation humidity of climt.EmanuelConvection)
//   compute_qs / _sat_vap_pressure  climt/_core/condensibles.py:66-76, 104-123  (of climt.EmanuelConvectionPython)
// for IPBL = 0 (the reference's shim forces it, _emanuel_convection.pyx:71-73) and NTRA = 0 (component.py:228).
//
// One WARP = one column.  The routine is a chain of short serial scans along the levels (geopotential, CAPE, the
// precipitating downdraft) between blocks of work that are independent per level or per (origin, destination) level pair;
// a first version with one thread per column ran 4 of 32 lanes on average (every column has its own cloud base, cloud top
// and branches) behind a 30 M-cycle dependent chain of global loads (r01 ncu: 47 ms for 64 800 columns, 15 ms for 8 100).
// Here
//   * the ~30 per-level vectors of a column live in the warp's slice of shared memory,
//   * level-parallel blocks (saturation humidity, TLIFT's Newton iterations, the per-origin-level normalisation, the per-level
//     detrainment sums and tendencies) give one level to each lane; the (origin I, destination J) mixing block runs I serially
//     with J across the lanes,
//   * the serial scans are executed redundantly by all 32 lanes with their recurrences in registers (only values that the
//     section does not itself modify are read from shared memory): no broadcast, no divergence, and the sums keep the
//     Fortran's order, so that results agree with the serial code to the last bits,
//   * the six (level x level) mixing matrices (MENT, QENT, ELIJ, SIJ, UENT, VENT) live in a per-warp-slot global workspace,
//     destination level fastest; a persistent grid reuses a slot for the next column, so the workspace stays L2-resident.
// A column is contiguous in the component's (column, level) layout and strided in the radiation engines' (level, column)
// layout; both are read in place through (level stride, column stride).
//
// What differs from a transliteration of the Fortran (results unchanged: only exact zeros are skipped):
//  * only the rectangle of the matrices that is ever read (rows ICB+1..INB x columns ICB-1..INB+1) is initialised, after INB is
//    known -- the Fortran initialises all six (NL+1)^2 matrices for every column (:594-610, 160 KB per column at 60 levels);
//  * MENT is non-zero only inside rows ICB+1..INB x columns ICB..INB: the O(n^3) mass-flux sums (:1018-1027), the detrainment
//    sums (:869-875, :1055-1074) and the level-1 entrainment loop (:995-1003, column 1 < ICB: identically zero) run over that
//    rectangle only, in the Fortran's own order of accumulation -- except the two O(n^3) mass-flux sums, which are formed from
//    cumulative column sums of MENT (O(n^2); last-bit differences);
//  * SIGP(I) = SIGS for every level (:567, :581) is a scalar; TH (:327-330) is computed but never used and is dropped.
#pragma once
#include <math.h>
#include <stdint.h>

#include "cb_common.h"

#if defined(__CUDA_ARCH__)
#define CB_LANES_FOR(i, lo, hi) _Pragma("unroll 1") for (int i = (lo) + lane; i <= (hi); i += 32)
#define CB_WARP_SYNC() __syncwarp()
#define CB_WARP_SUM_INT(x) __reduce_add_sync(0xffffffffu, (x))
#define CB_WARP_ANY(x) (__any_sync(0xffffffffu, (x)) != 0)
#else
#define CB_LANES_FOR(i, lo, hi) for (int i = (lo); i <= (hi); ++i)
#define CB_WARP_SYNC()
#define CB_WARP_SUM_INT(x) (x)
#define CB_WARP_ANY(x) (x)
#endif
// Out-of-line on the device: the Newton step of TLIFT carries two exp and one log expansion per iteration; inlined at both call
// sites and unrolled it was a fifth of the kernel's 12 168 instructions (r01 ncu: 69 % of the stall samples were instruction fetch).
#if defined(__CUDACC__)
#define CB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CB_HD_NOINLINE inline
#endif

namespace cb {
namespace emanuel {

// same member order as cb200_emanuel_params (include/climt_b200.h)
struct Par {
  double minorig, elcrit, tlcrit, entp, sigd, sigs, omtrain, omtsnow, coeffr, coeffs, cu, beta, dtmax, alpha, damp;
  double cpd, cpv, cl, rv, rd, lv0, g, rowl, delt0, t_rain;
};

enum QsMode { QS_GIVEN = 0, QS_BOLTON = 1, QS_PYTHON = 2 };

// device pointers; element (level k, column c) of a 2-D array is p[k * ls + c * cs]; level 0 at the surface, pressures in mbar
struct In {
  int nlev;
  size_t ls, cs;       // level / column strides of the (nlev) arrays
  size_t ls_i, cs_i;   // of ph (nlev + 1 levels)
  const double *t, *q, *u, *v, *p, *ph;
  const double* qs;    // QS_GIVEN only
  const double* cbmf;  // (ncol) cloud-base mass flux of the previous step
  int qs_mode;
};

struct Out {
  size_t ls, cs;
  double *ft, *fq, *fu, *fv;                            // tendencies [K s-1, kg kg-1 s-1, m s-2]
  double *precip, *wd, *tprime, *qprime, *cbmf, *cape;  // (ncol)
  int32_t* iflag;                                       // (ncol)
};

// per-level vectors of one column in shared memory (1-based index, n1 = nlev + 4 entries each)
enum Vec { V_T = 0, V_Q, V_U, V_V, V_P, V_PH, V_QS, V_M, V_MP, V_TVP, V_TV, V_WATER, V_QP, V_EP, V_WT, V_EVAP, V_CLW, V_TP, V_CPN,
           V_LV, V_LVCP, V_H, V_HP, V_GZ, V_HM, V_UP, V_VP, V_NENT, V_WDT, V_FT, V_FQ, V_FU, V_FV, V_COUNT };
enum Mat { M_MENT = 0, M_QENT, M_ELIJ, M_SIJ, M_UENT, M_VENT, M_COUNT };

struct Work {
  int n1;     // entries per shared-memory vector
  int nm;     // rows = columns of a mixing matrix (NL + 2): index 0..NL+1
  double* m;  // [slot][Mat][nm][nm], destination level fastest
};

struct SV {  // 1-based vector
  double* p;
  CB_HD double& operator()(int i) const { return p[i]; }
};
struct M2 {
  double* p;
  int n;
  CB_HD double& operator()(int i, int j) const { return p[(size_t)i * n + j]; }
};

// saturation vapour pressure over water / ice in hPa (TLIFT, convect43c.f90:1203-1208 = condensibles.py:72-76)
CB_HD double es_hpa(double TG) {
  const double TC = TG - 273.15;
  if (TC >= 0.0) return 6.112 * exp(17.67 * TC / (243.5 + TC));
  return exp(23.33086 - 6111.72784 / TG + 0.15215 * log(TG));
}

CB_HD double saturation_q(const Par& c, int mode, double T, double p_mbar) {
  const double eps = c.rd / c.rv;
  if (mode == QS_BOLTON) {  // util.py:177-180 with p in Pa
    const double es = 611.2 * exp(17.67 * (T - 273.15) / (T - 29.65));
    return eps * es / (p_mbar * 100 - (1 - eps) * es);
  }
  const double es = es_hpa(T);  // condensibles.py:120-122
  return eps * es / (p_mbar - (1.0 - eps) * es);
}

// One level of TLIFT above cloud base (convect43c.f90:1192-1217): two Newton steps on the saturated parcel's temperature
CB_HD_NOINLINE void tlift_level(const Par& c, double AH0, double qnk, double ti, double pi, double gz, double qsi, double& tpk, double& clw,
                       double& tvp) {
  const double CPVMCL = c.cl - c.cpv, EPS = c.rd / c.rv, EPSI = 1. / EPS;
  double TG = ti, QG = qsi;
  const double ALV = c.lv0 - CPVMCL * (ti - 273.15);
#pragma unroll 1
  for (int J = 1; J <= 2; ++J) {
    const double S = 1. / (c.cpd + ALV * ALV * QG / (c.rv * ti * ti));
    const double AHG = c.cpd * TG + (c.cl - c.cpd) * qnk * ti + ALV * QG + gz;
    TG = fmax(TG + S * (AH0 - AHG), 35.0);
    const double ES = es_hpa(TG);
    QG = EPS * ES / (pi - ES * (1. - EPS));
  }
  tpk = (AH0 - (c.cl - c.cpd) * qnk * ti - gz - ALV * QG) / c.cpd;
  clw = fmax(0.0, qnk - QG);
  const double RG = QG / (1. - qnk);
  tvp = tpk * (1. + RG * EPSI);
}

// CONVECT (convect43c.f90:146-1148) for column `col`, executed by one warp (`lane` = 0..31; the host build runs it as one lane).
// sv: the warp's V_COUNT * W.n1 doubles of shared memory; mw: the warp slot's M_COUNT * nm * nm doubles.  ND = nlev, NL = max_conv_lev.
CB_HD void convect_warp(const Par& c, const In& in, const Work& W, const Out& out, size_t col, int lane, double* sv, double* mw, int NL,
                        double DELT) {
  (void)lane;
  const int ND = in.nlev;
  auto vec = [&](int f) { return SV{sv + (size_t)f * W.n1}; };
  auto mat = [&](int f) { return M2{mw + (size_t)f * W.nm * W.nm, W.nm}; };
  const SV T = vec(V_T), Q = vec(V_Q), U = vec(V_U), V = vec(V_V), P = vec(V_P), PH = vec(V_PH), QS = vec(V_QS), M = vec(V_M),
           MP = vec(V_MP), TVP = vec(V_TVP), TV = vec(V_TV), WATER = vec(V_WATER), QP = vec(V_QP), EP = vec(V_EP), WT = vec(V_WT),
           EVAP = vec(V_EVAP), CLW = vec(V_CLW), TP = vec(V_TP), CPN = vec(V_CPN), LV = vec(V_LV), LVCP = vec(V_LVCP), H = vec(V_H),
           HP = vec(V_HP), GZ = vec(V_GZ), HM = vec(V_HM), UP = vec(V_UP), VP = vec(V_VP), NENT = vec(V_NENT), WDT = vec(V_WDT),
           FT = vec(V_FT), FQ = vec(V_FQ), FU = vec(V_FU), FV = vec(V_FV);
  const M2 MENT = mat(M_MENT), QENT = mat(M_QENT), ELIJ = mat(M_ELIJ), SIJ = mat(M_SIJ), UENT = mat(M_UENT), VENT = mat(M_VENT);
  const double CPVMCL = c.cl - c.cpv, EPS = c.rd / c.rv, EPSI = 1. / EPS, GINV = 1.0 / c.g, DELTI = 1.0 / DELT;
  const int MINORIG = (int)c.minorig;

  // ---- the column into shared memory; saturation humidity (the components' host-side pre-step, fused); zeroed tendencies
  CB_WARP_SYNC();  // the previous column of this warp is done with the vectors
  {
    const double *pt = in.t + col * in.cs, *pq = in.q + col * in.cs, *pu = in.u + col * in.cs, *pv = in.v + col * in.cs,
                 *pp = in.p + col * in.cs, *pph = in.ph + col * in.cs_i;
    CB_LANES_FOR(I, 1, ND + 1) {
      PH(I) = pph[(size_t)(I - 1) * in.ls_i];
      if (I <= ND) {
        const size_t o = (size_t)(I - 1) * in.ls;
        const double t = pt[o], p = pp[o];
        T(I) = t; Q(I) = pq[o]; U(I) = pu[o]; V(I) = pv[o]; P(I) = p;
        QS(I) = in.qs_mode == QS_GIVEN ? in.qs[col * in.cs + o] : saturation_q(c, in.qs_mode, t, p);
        FT(I) = 0.0; FQ(I) = 0.0; FU(I) = 0.0; FV(I) = 0.0;
      }
    }
  }
  CB_WARP_SYNC();
  double CBMF = in.cbmf[col];
  int IFLAG = 0;
  double PRECIP = 0.0, WD = 0.0, TPRIME = 0.0, QPRIME = 0.0, OUTCAPE = 0.0;
  auto finish = [&]() {
    CB_WARP_SYNC();
    double *oft = out.ft + col * out.cs, *ofq = out.fq + col * out.cs, *ofu = out.fu + col * out.cs, *ofv = out.fv + col * out.cs;
    CB_LANES_FOR(I, 1, ND) {
      const size_t o = (size_t)(I - 1) * out.ls;
      oft[o] = FT(I); ofq[o] = FQ(I); ofu[o] = FU(I); ofv[o] = FV(I);
    }
    if (lane == 0) {
      out.iflag[col] = IFLAG; out.precip[col] = PRECIP; out.wd[col] = WD; out.tprime[col] = TPRIME; out.qprime[col] = QPRIME;
      out.cbmf[col] = CBMF; out.cape[col] = OUTCAPE;
    }
  };

  // ---- geopotential, heat capacity, static energies; level of minimum moist static energy (:426-458).  Serial scan, all lanes.
  int IHMIN = NL;
  const double t1 = T(1), q1 = Q(1);
  {
    const double cpn1 = c.cpd * (1. - q1) + q1 * c.cpv, lv1 = c.lv0 - CPVMCL * (t1 - 273.15);
    double hm_prev = lv1 * q1, tv_prev = t1 * (1. + q1 * EPSI - q1), gz = 0.0, p_prev = P(1), AHMIN = 1.0E12;
    GZ(1) = 0.0; CPN(1) = cpn1; H(1) = t1 * cpn1; LV(1) = lv1; HM(1) = hm_prev; TV(1) = tv_prev;
#pragma unroll 1
    for (int I = 2; I <= NL + 1; ++I) {
      const double ti = T(I), qi = Q(I), pi = P(I);
      const double tvx = ti * (1. + qi * EPSI - qi);
      gz = gz + 0.5 * c.rd * (tvx + tv_prev) * (p_prev - pi) / PH(I);
      const double cpn = c.cpd * (1. - qi) + c.cpv * qi, lv = c.lv0 - CPVMCL * (ti - 273.15);
      const double hm = (c.cpd * (1. - qi) + c.cl * qi) * (ti - t1) + lv * qi + gz;
      GZ(I) = gz; CPN(I) = cpn; H(I) = ti * cpn + gz; LV(I) = lv; HM(I) = hm; TV(I) = tvx;
      if (I >= MINORIG && hm < AHMIN && hm < hm_prev) { AHMIN = hm; IHMIN = I; }
      hm_prev = hm; tv_prev = tvx; p_prev = pi;
    }
  }
  IHMIN = IHMIN < NL - 1 ? IHMIN : NL - 1;
  CB_WARP_SYNC();
  // level of maximum moist static energy at or below IHMIN (:463-470).  NK stays at the first level when no HM is positive
  // (the Fortran would index T(0); the reference's own port reads its level 0, pure_python_v3.py:337-343).
  int NK = 1;
  {
    double AHMAX = 0.0;
#pragma unroll 1
    for (int I = MINORIG; I <= IHMIN; ++I) {
      const double hm = HM(I);
      if (hm > AHMAX) { NK = I; AHMAX = hm; }
    }
  }
  const double tnk = T(NK), qnk = Q(NK), unk = U(NK), vnk = V(NK);
  if (tnk < 250.0 || qnk <= 0.0 || IHMIN == (NL - 1)) { IFLAG = 0; CBMF = 0.0; finish(); return; }
  // lifted condensation level (:491-503) and first level above it (:509-519)
  const double RH = qnk / QS(NK);
  const double CHI = tnk / (1669.0 - 122.0 * RH - tnk);
  const double PLCL = P(NK) * pow(RH, CHI);
  if (PLCL < 200.0 || PLCL >= 2000.0) { IFLAG = 2; CBMF = 0.0; finish(); return; }
  int ICB = NL - 1;
#pragma unroll 1
  for (int I = NK + 1; I <= NL; ++I)
    if (P(I) < PLCL) { ICB = ICB < I ? ICB : I; }
  if (ICB >= (NL - 1)) { IFLAG = 3; CBMF = 0.0; finish(); return; }
  // ---- parcel up to ICB (TLIFT with KK = 1, :1169-1217) and the stability test (:528-540)
  const double AH0 = (c.cpd * (1. - qnk) + c.cl * qnk) * tnk + qnk * (c.lv0 - CPVMCL * (tnk - 273.15)) + GZ(NK);
  {
    const double CPINV = 1. / (c.cpd * (1. - qnk) + qnk * c.cpv), gznk = GZ(NK);
    CB_LANES_FOR(I, 1, ICB - 1) {
      CLW(I) = 0.0;
      if (I >= NK) {
        const double tpk = tnk - (GZ(I) - gznk) * CPINV;
        TP(I) = tpk;
        TVP(I) = tpk * (1. + qnk * EPSI) - tpk * qnk;  // TVP(I) - TP(I)*Q(NK) of :529-531 applied at once
      }
    }
  }
  double tvp_icb;
  {
    double tpk, clw, tvp;
    tlift_level(c, AH0, qnk, T(ICB), P(ICB), GZ(ICB), QS(ICB), tpk, clw, tvp);
    tvp_icb = tvp - tpk * qnk;
    TP(ICB) = tpk; CLW(ICB) = clw; TVP(ICB) = tvp_icb;
  }
  if (CBMF == 0.0 && tvp_icb <= (TV(ICB) - c.dtmax)) { IFLAG = 0; finish(); return; }
  IFLAG = 1;
  // ---- the rest of the parcel (TLIFT with KK = 2), one level per lane, with :588-590 applied at once
  CB_LANES_FOR(I, ICB + 1, NL) {
    double tpk, clw, tvp;
    tlift_level(c, AH0, qnk, T(I), P(I), GZ(I), QS(I), tpk, clw, tvp);
    TP(I) = tpk; CLW(I) = clw; TVP(I) = tvp - tpk * qnk;
  }
  CB_WARP_SYNC();
  // precipitation efficiencies (:564-582) and the per-level work vectors (:594-626)
  CB_LANES_FOR(I, 1, NL + 1) {
    if (I <= NK) {
      EP(I) = 0.0;
    } else if (I <= NL) {
      const double TCA = TP(I) - 273.15;
      double ELACRIT = TCA >= 0.0 ? c.elcrit : c.elcrit * (1.0 - TCA / c.tlcrit);
      ELACRIT = fmax(ELACRIT, 0.0);
      const double EPMAX = 0.999;
      double ep = EPMAX * (1.0 - ELACRIT / fmax(CLW(I), 1.0E-8));
      ep = fmax(ep, 0.0);
      EP(I) = fmin(ep, EPMAX);
    }
    HP(I) = H(I); NENT(I) = 0.0; WATER(I) = 0.0; EVAP(I) = 0.0; WT(I) = c.omtsnow; MP(I) = 0.0; M(I) = 0.0;
    LVCP(I) = LV(I) / CPN(I);
    QP(I) = I == 1 ? Q(1) : Q(I - 1); UP(I) = I == 1 ? U(1) : U(I - 1); VP(I) = I == 1 ? V(1) : V(I - 1);
    if (I == NL + 1) TVP(NL + 1) = TVP(NL) - (GZ(NL + 1) - GZ(NL)) / c.cpd;
  }
  CB_WARP_SYNC();
  // ---- level of neutral buoyancy INB and CAPE (:632-655).  Serial scan, all lanes.
  int INB = ICB + 1, INB1 = INB;
  double FRAC;
  {
    double CAPE = 0.0, CAPEM = 0.0, BYP = 0.0;
#pragma unroll 1
    for (int I = ICB + 1; I <= NL - 1; ++I) {
      const double BY = (TVP(I) - TV(I)) * (PH(I) - PH(I + 1)) / P(I);
      CAPE = CAPE + BY;
      if (BY >= 0.0) INB1 = I + 1;
      if (CAPE > 0.0) {
        INB = I + 1;
        BYP = (TVP(I + 1) - TV(I + 1)) * (PH(I + 1) - PH(I + 2)) / P(I + 1);
        CAPEM = CAPE;
      }
    }
    INB = INB > INB1 ? INB : INB1;
    CAPE = CAPEM + BYP;
    const double DEFRAC = fmax(CAPEM - CAPE, 0.001);
    FRAC = fmax(fmin(-CAPE / DEFRAC, 1.0), 0.0);
    OUTCAPE = CAPE;
  }
  // ---- cloud-base mass flux (:667-704)
  {
    const double tvpm = TVP(ICB - 1), pm = P(ICB - 1);
    const double TVPPLCL = tvpm - c.rd * tvpm * (pm - PLCL) / (CPN(ICB - 1) * pm);
    const double TVAPLCL = TV(ICB) + (TVP(ICB) - TVP(ICB + 1)) * (PLCL - P(ICB)) / (P(ICB) - P(ICB + 1));
    double DTPBL = 0.0;
#pragma unroll 1
    for (int I = NK; I <= ICB - 1; ++I) DTPBL = DTPBL + (TVP(I) - TV(I)) * (PH(I) - PH(I + 1));
    DTPBL = DTPBL / (PH(NK) - PH(ICB));
    const double DTMA = TVPPLCL - TVAPLCL + c.dtmax + DTPBL;
    const double CBMFOLD = CBMF;
    const double DAMPS = c.damp * DELT / c.delt0;
    CBMF = (1. - DAMPS) * CBMF + 0.1 * c.alpha * DTMA;
    CBMF = fmax(CBMF, 0.0);
    if (CBMF == 0.0 && CBMFOLD == 0.0) { finish(); return; }
  }
  // ---- liquid water static energy of the lifted parcel (:660-662), rates of mixing M(I) (:708-718; the sum in the Fortran's order),
  //      and the part of the mixing matrices that is ever read: rows ICB+1..INB, columns ICB-1..INB+1 (defaults of :602-607)
  {
    double DBOSUM = 0.0;
#pragma unroll 1
    for (int I = ICB + 1; I <= INB; ++I) {
      const int K = I < INB1 ? I : INB1;
      DBOSUM = DBOSUM + (fabs(TV(K) - TVP(K)) + c.entp * 0.02 * (PH(K) - PH(K + 1)));
    }
    const double hnk = H(NK);
    CB_LANES_FOR(I, ICB, INB) {
      HP(I) = hnk + (LV(I) + (c.cpd - c.cpv) * T(I)) * EP(I) * CLW(I);
      if (I > ICB) {
        const int K = I < INB1 ? I : INB1;
        const double DBO = fabs(TV(K) - TVP(K)) + c.entp * 0.02 * (PH(K) - PH(K + 1));
        M(I) = (CBMF * DBO) / DBOSUM;
      }
    }
#pragma unroll 1
    for (int I = ICB + 1; I <= INB; ++I)
      CB_LANES_FOR(J, ICB - 1, INB + 1) {
        MENT(I, J) = 0.0; ELIJ(I, J) = 0.0; SIJ(I, J) = 0.0;
        QENT(I, J) = Q(J); UENT(I, J) = U(J); VENT(I, J) = V(J);
      }
  }
  CB_WARP_SYNC();
  // ---- entrained mass flux, total water, condensed water and mixing fraction of every (origin I, destination J) pair (:727-786):
  //      origin levels in turn, destination levels across the lanes
#pragma unroll 1
  for (int I = ICB + 1; I <= INB; ++I) {
    const double qi = Q(I), ui = U(I), vi = V(I), hi = H(I), hpi = HP(I), mi = M(I);
    const double QTI = qnk - EP(I) * CLW(I);
    int nent = 0;
    CB_LANES_FOR(J, ICB, INB) {
      const double tj = T(J), qsj = QS(J), lvj = LV(J);
      const double BF2 = 1. + lvj * lvj * qsj / (c.rv * tj * tj * c.cpd);
      double ANUM = H(J) - hpi + (c.cpv - c.cpd) * tj * (QTI - Q(J));
      double DENOM = hi - hpi + (c.cpd - c.cpv) * (qi - QTI) * tj;
      double DEI = DENOM;
      if (fabs(DEI) < 0.01) DEI = 0.01;
      double sij = J == I ? 1.0 : ANUM / DEI;  // SIJ(I,I)=1.0 is stored before SIJ(I,J) is read back (:735-736)
      double ALTEM = (sij * qi + (1. - sij) * QTI - qsj) / BF2;
      const double CWAT = CLW(J) * (1. - EP(J));
      if ((sij < 0.0 || sij > 1.0 || ALTEM > CWAT) && J > I) {
        ANUM = ANUM - lvj * (QTI - qsj - CWAT * BF2);
        DENOM = DENOM + lvj * (qi - QTI);
        if (fabs(DENOM) < 0.01) DENOM = 0.01;
        sij = ANUM / DENOM;
        ALTEM = sij * qi + (1. - sij) * QTI - qsj;
        ALTEM = ALTEM - (BF2 - 1.) * CWAT;
      }
      if (sij > 0.0 && sij < 0.9) {
        QENT(I, J) = sij * qi + (1. - sij) * QTI;
        UENT(I, J) = sij * ui + (1. - sij) * unk;
        VENT(I, J) = sij * vi + (1. - sij) * vnk;
        ELIJ(I, J) = fmax(0.0, ALTEM);
        MENT(I, J) = mi / (1. - sij);
        ++nent;
      }
      SIJ(I, J) = fmin(1.0, fmax(0.0, sij));
    }
    nent = CB_WARP_SUM_INT(nent);
    // no air can entrain at level I: the updraft detrains there (:776-785)
    if (lane == 0) {
      if (nent == 0) { MENT(I, I) = mi; QENT(I, I) = QTI; UENT(I, I) = unk; VENT(I, I) = vnk; ELIJ(I, I) = CLW(I); SIJ(I, I) = 1.0; }
      if (I == INB) SIJ(INB, INB) = 1.0;
      NENT(I) = (double)nent;
    }
  }
  CB_WARP_SYNC();
  // ---- normalise the entrained fluxes to equal probabilities of mixing (:792-856): one origin level per lane
  CB_LANES_FOR(I, ICB + 1, INB) {
    if (NENT(I) != 0.0) {
      const double qi = Q(I), hi = H(I), hpi = HP(I), lvi = LV(I), qsi = QS(I);
      const double QP1 = qnk - EP(I) * CLW(I);
      const double ANUM = hi - hpi - lvi * (QP1 - qsi);
      double DENOM = hi - hpi + lvi * (qi - QP1);
      if (fabs(DENOM) < 0.01) DENOM = 0.01;
      double SCRIT = ANUM / DENOM;
      const double ALT = QP1 - qsi + SCRIT * (qi - QP1);
      if (ALT < 0.0) SCRIT = 1.0;
      SCRIT = fmax(SCRIT, 0.0);
      double ASIJ = 0.0, SMIN = 1.0;
      double s_lo = SIJ(I, ICB - 1), s_j = SIJ(I, ICB);  // sliding window SIJ(I,J-1), SIJ(I,J), SIJ(I,J+1)
#pragma unroll 1
      for (int J = ICB; J <= INB; ++J) {
        const double s_hi = SIJ(I, J + 1);
        if (s_j > 0.0 && s_j < 0.9) {
          double SMID, SJMAX, SJMIN;
          if (J > I) {
            SMID = fmin(s_j, SCRIT);
            SJMAX = SMID;
            SJMIN = SMID;
            if (SMID < SMIN && s_hi < SMID) {
              SMIN = SMID;
              SJMAX = fmin(fmin(s_hi, s_j), SCRIT);
              SJMIN = fmin(fmax(s_lo, s_j), SCRIT);
            }
          } else {
            SJMAX = fmax(s_hi, SCRIT);
            SMID = fmax(s_j, SCRIT);
            SJMIN = fmax(J > 1 ? s_lo : 0.0, SCRIT);
          }
          const double DELP = fabs(SJMAX - SMID), DELM = fabs(SJMIN - SMID);
          const double dph = PH(J) - PH(J + 1);
          ASIJ = ASIJ + (DELP + DELM) * dph;
          MENT(I, J) = MENT(I, J) * (DELP + DELM) * dph;
        }
        s_lo = s_j;
        s_j = s_hi;
      }
      ASIJ = 1.0 / fmax(1.0E-21, ASIJ);
      double BSUM = 0.0;
#pragma unroll 1
      for (int J = ICB; J <= INB; ++J) {
        const double m = MENT(I, J) * ASIJ;
        MENT(I, J) = m;
        BSUM = BSUM + m;
      }
      if (BSUM < 1.0E-18) {
        NENT(I) = 0.0;
        MENT(I, I) = M(I); QENT(I, I) = QP1; UENT(I, I) = unk; VENT(I, I) = vnk; ELIJ(I, I) = CLW(I); SIJ(I, I) = 1.0;
      }
    }
  }
  CB_WARP_SYNC();
  // ---- precipitating downdraft (:866-956)
  if (!(EP(INB) < 0.0001)) {
    // detrained precipitation of every level (:869-875), one level per lane; MENT(J,I) = 0 for destination levels below cloud base
    CB_LANES_FOR(I, 1, INB) {
      const double epi = EP(I), clwi = CLW(I);
      double WDTRAIN = c.g * epi * M(I) * clwi;
      if (I >= ICB)
#pragma unroll 1
        for (int J = ICB + 1; J <= I - 1; ++J) {
          const double AWAT = fmax(0.0, ELIJ(J, I) - (1. - epi) * clwi);
          WDTRAIN = WDTRAIN + c.g * AWAT * MENT(J, I);
        }
      WDT(I) = WDTRAIN;
    }
    CB_WARP_SYNC();
    // condensed water, evaporation, downdraft mass flux, humidity and momentum: serial from the top down, all lanes, the
    // recurrence (level I+1's WATER, WT, MP, QP, UP, VP) in registers; only stores to the vectors it produces
    int JTT = 2;
    const double p1 = P(1);
    double water1 = 0.0, wt1 = c.omtsnow, mp1 = 0.0, qp1 = Q(INB), up1 = U(INB), vp1 = V(INB);  // level INB+1 (:594-626)
    double mp_jtt = 0.0, p_jtt = P(2);
#pragma unroll 1
    for (int I = INB; I >= 1; --I) {
      const double ti = T(I), qi = Q(I), qsi = QS(I), phi = PH(I), phi1 = PH(I + 1);
      double COEFF = c.coeffs, wt = c.omtsnow;
      if (ti > c.t_rain) { COEFF = c.coeffr; wt = c.omtrain; }
      const double QSM = 0.5 * (qi + qp1);
      const double AFAC = fmax(COEFF * phi * (qsi - QSM) / (1.0E4 + 2.0E3 * phi * qsi), 0.0);
      const double SIGT = fmin(1.0, fmax(0.0, c.sigs));
      const double B6 = 100. * (phi - phi1) * SIGT * AFAC / wt;
      const double C6 = (water1 * wt1 + WDT(I) / c.sigd) / wt;
      const double REVAP = 0.5 * (-B6 + sqrt(B6 * B6 + 4. * C6));
      const double evap = SIGT * AFAC * REVAP;
      const double water = REVAP * REVAP;
      double mp = 0.0;  // MP(1) keeps its initial zero (:901)
      if (I != 1) {
        const double DHDP = fmax((H(I) - H(I - 1)) / (P(I - 1) - P(I)), 10.0);
        mp = fmax(100. * GINV * LV(I) * c.sigd * evap / DHDP, 0.0);
        const double FAC = 20.0 / (PH(I - 1) - phi);
        mp = (FAC * mp1 + mp) / (1. + FAC);
        const double pi = P(I);
        if (pi > (0.949 * p1)) {  // force MP to decrease linearly to zero between about 950 mb and the surface (:913-919)
          if (I >= JTT) { JTT = I; mp_jtt = mp; p_jtt = pi; }   // JTT = MAX(JTT, I); MP(JTT) is read after MP(I) was stored
          mp = mp_jtt * (p1 - pi) / (p1 - p_jtt);
          if (JTT == I) mp_jtt = mp;
        }
      }
      // QP / UP / VP of this level: initial values (:611-626) unless the downdraft branches below replace them
      double qp = I == 1 ? Q(1) : Q(I - 1), up = I == 1 ? U(1) : U(I - 1), vp = I == 1 ? V(1) : V(I - 1);
      if (I != INB) {
        const double QSTM = I == 1 ? QS(1) : QS(I - 1);
        if (mp > mp1) {
          const double RAT = mp1 / mp;
          qp = qp1 * RAT + qi * (1.0 - RAT) + 100. * GINV * c.sigd * (phi - phi1) * (evap / mp);
          up = up1 * RAT + U(I) * (1. - RAT);
          vp = vp1 * RAT + V(I) * (1. - RAT);
        } else if (mp1 > 0.0) {
          qp = (GZ(I + 1) - GZ(I) + qp1 * (LV(I + 1) + T(I + 1) * (c.cl - c.cpd)) + c.cpd * (T(I + 1) - ti)) / (LV(I) + ti * (c.cl - c.cpd));
          up = up1;
          vp = vp1;
        }
        qp = fmax(fmin(qp, QSTM), 0.0);
      }
      EVAP(I) = evap; WATER(I) = water; WT(I) = wt; MP(I) = mp; QP(I) = qp; UP(I) = up; VP(I) = vp;
      water1 = water; wt1 = wt; mp1 = mp; qp1 = qp; up1 = up; vp1 = vp;
    }
    PRECIP = PRECIP + wt1 * c.sigd * water1 * 3600. * 24000. / (c.rowl * c.g);  // WT(1), WATER(1)
    CB_WARP_SYNC();
  }
  // downdraft velocity scale and surface fluctuations (:966-968)
  WD = c.beta * fabs(MP(ICB)) * 0.01 * c.rd * T(ICB) / (c.sigd * P(ICB));
  QPRIME = 0.5 * (QP(1) - Q(1));
  TPRIME = c.lv0 * QPRIME / c.cpd;
  // ---- cumulative mass flux entrained into each destination level from origins at or below K:
  //      CS(K, J) = sum over K' = ICB+1..K of MENT(K', J), stored over SIJ (no longer needed); destination levels across the lanes.
  //      The net saturated up- and downdraft fluxes through a level (:1018-1030) are sums of O(n) of these instead of O(n^2)
  //      entries of MENT each -- the one place where the order of a summation differs from the Fortran's (last-bit differences).
  const M2 CS = SIJ;
  CB_LANES_FOR(J, ICB, INB) {
    double cs = 0.0;
#pragma unroll 4
    for (int K = ICB + 1; K <= INB; ++K) {
      cs = cs + MENT(K, J);
      CS(K, J) = cs;
    }
  }
  CB_WARP_SYNC();
  // ---- tendencies: one level per lane.  Level 1 (:974-994; its entrainment loop :995-1003 adds exact zeros: column 1 < ICB),
  //      levels 2..INB (:1012-1086)
  bool cfl = false;
  CB_LANES_FOR(I, 1, INB) {
    const double DPINV = 0.01 / (PH(I) - PH(I + 1));
    if (I == 1) {
      double AM = 0.0;
      if (NK == 1)
#pragma unroll 1
        for (int K = 2; K <= INB; ++K) AM = AM + M(K);
      if ((2. * c.g * DPINV * AM) >= DELTI) cfl = true;
      const double t2 = T(2), u1 = U(1), v1 = V(1), cpn1 = CPN(1), evap1 = EVAP(1), mp2 = MP(2);
      double ft = c.g * DPINV * AM * (t2 - t1 + (GZ(2) - GZ(1)) / cpn1);
      ft = ft - LVCP(1) * c.sigd * evap1;
      ft = ft + c.sigd * WT(2) * (c.cl - c.cpd) * WATER(2) * (t2 - t1) * DPINV / cpn1;
      FT(1) = ft;
      double fq = c.g * mp2 * (QP(2) - q1) * DPINV + c.sigd * evap1;
      fq = fq + c.g * AM * (Q(2) - q1) * DPINV;
      FQ(1) = fq;
      FU(1) = c.g * DPINV * (mp2 * (UP(2) - u1) + AM * (U(2) - u1));
      FV(1) = c.g * DPINV * (mp2 * (VP(2) - v1) + AM * (V(2) - v1));
    } else {
      const double CPINV = 1.0 / CPN(I);
      // net saturated updraft (AMP1) and downdraft (AD) mass fluxes through level I (:1018-1030)
      double AMP1 = 0.0, AD = 0.0;
      if (I >= NK)
#pragma unroll 1
        for (int K = I + 1; K <= INB + 1; ++K) AMP1 = AMP1 + M(K);
      if (I >= ICB + 1) {  // air from origins ICB+1..I that mixes into levels above I: sum over J = I+1..INB of CS(min(I, INB), J)
        const int k_hi = I < INB ? I : INB;
#pragma unroll 4
        for (int J = I + 1; J <= INB; ++J) AMP1 = AMP1 + CS(k_hi, J);
      }
      if ((2. * c.g * DPINV * AMP1) >= DELTI) cfl = true;
      if (I >= ICB + 1) {  // air from origins I..INB that mixes into levels K = ICB..I-1: column totals minus the part from below I
        const bool below = I - 1 >= ICB + 1;
#pragma unroll 4
        for (int K = ICB; K <= I - 1; ++K) AD = AD + (CS(INB, K) - (below ? CS(I - 1, K) : 0.0));
      }
      const double ti = T(I), qi = Q(I), ui = U(I), vi = V(I), tm = T(I - 1), tp = T(I + 1), gzi = GZ(I), evap = EVAP(I);
      double ft = c.g * DPINV * (AMP1 * (tp - ti + (GZ(I + 1) - gzi) * CPINV) - AD * (ti - tm + (gzi - GZ(I - 1)) * CPINV)) -
                  c.sigd * LVCP(I) * evap;
      if (I >= ICB + 1) ft = ft + c.g * DPINV * MENT(I, I) * (HP(I) - H(I) + ti * (c.cpv - c.cpd) * (qi - QENT(I, I))) * CPINV;
      ft = ft + c.sigd * WT(I + 1) * (c.cl - c.cpd) * WATER(I + 1) * (tp - ti) * DPINV * CPINV;
      FT(I) = ft;
      double fq = c.g * DPINV * (AMP1 * (Q(I + 1) - qi) - AD * (qi - Q(I - 1)));
      double fu = c.g * DPINV * (AMP1 * (U(I + 1) - ui) - AD * (ui - U(I - 1)));
      double fv = c.g * DPINV * (AMP1 * (V(I + 1) - vi) - AD * (vi - V(I - 1)));
      if (I >= ICB) {
        const double wat_i = (1. - EP(I)) * CLW(I);
        const int k_mid = I - 1 < INB ? I - 1 : INB;
#pragma unroll 1
        for (int K = ICB + 1; K <= k_mid; ++K) {  // air detrained from below (:1055-1064)
          const double AWAT = fmax(ELIJ(K, I) - wat_i, 0.0), ment = MENT(K, I);
          fq = fq + c.g * DPINV * ment * (QENT(K, I) - AWAT - qi);
          fu = fu + c.g * DPINV * ment * (UENT(K, I) - ui);
          fv = fv + c.g * DPINV * ment * (VENT(K, I) - vi);
        }
#pragma unroll 1
        for (int K = I > ICB + 1 ? I : ICB + 1; K <= INB; ++K) {  // from this level and above (:1065-1074)
          const double ment = MENT(K, I);
          fq = fq + c.g * DPINV * ment * (QENT(K, I) - qi);
          fu = fu + c.g * DPINV * ment * (UENT(K, I) - ui);
          fv = fv + c.g * DPINV * ment * (VENT(K, I) - vi);
        }
      }
      const double mp1 = MP(I + 1), mp = MP(I);
      FQ(I) = fq + c.sigd * evap + c.g * (mp1 * (QP(I + 1) - qi) - mp * (QP(I) - Q(I - 1))) * DPINV;
      FU(I) = fu + c.g * (mp1 * (UP(I + 1) - ui) - mp * (UP(I) - U(I - 1))) * DPINV;
      FV(I) = fv + c.g * (mp1 * (VP(I + 1) - vi) - mp * (VP(I) - V(I - 1))) * DPINV;
    }
  }
  if (CB_WARP_ANY(cfl)) IFLAG = 4;
  CB_WARP_SYNC();
  // ---- move part of the top level's tendencies down to reflect the actual level of zero CAPE (:1092-1108)
  if (lane == 0) {
    const double r = (PH(INB) - PH(INB + 1)) / (PH(INB - 1) - PH(INB));
    const double fq = FQ(INB), ft = FT(INB), fu = FU(INB), fv = FV(INB);
    FQ(INB) = fq * (1. - FRAC);
    FQ(INB - 1) = FQ(INB - 1) + FRAC * fq * r * LV(INB) / LV(INB - 1);
    FT(INB) = ft * (1. - FRAC);
    FT(INB - 1) = FT(INB - 1) + FRAC * ft * r * CPN(INB) / CPN(INB - 1);
    FU(INB) = fu * (1. - FRAC);
    FU(INB - 1) = FU(INB - 1) + FRAC * fu * r;
    FV(INB) = fv * (1. - FRAC);
    FV(INB - 1) = FV(INB - 1) + FRAC * fv * r;
  }
  CB_WARP_SYNC();
  // ---- exact enthalpy and momentum conservation (:1122-1136): the integrals serially on all lanes, the correction per lane
  {
    double ENTS = 0.0, UAV = 0.0, VAV = 0.0;
#pragma unroll 1
    for (int I = 1; I <= INB; ++I) {
      const double dph = PH(I) - PH(I + 1);
      ENTS = ENTS + (CPN(I) * FT(I) + LV(I) * FQ(I)) * dph;
      UAV = UAV + FU(I) * dph;
      VAV = VAV + FV(I) * dph;
    }
    const double dp = PH(1) - PH(INB + 1);
    ENTS = ENTS / dp; UAV = UAV / dp; VAV = VAV / dp;
    CB_WARP_SYNC();
    CB_LANES_FOR(I, 1, INB) {
      FT(I) = FT(I) - ENTS / CPN(I);
      FU(I) = (1. - c.cu) * (FU(I) - UAV);
      FV(I) = (1. - c.cu) * (FV(I) - VAV);
    }
  }
  finish();
}

}  // namespace emanuel
}  // namespace cb
