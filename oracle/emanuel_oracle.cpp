// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use it.
//
// CPU restatement of the Emanuel convection scheme, following
//   SUBROUTINE CONVECT   climt/_lib/emanuel/convect43c.f90:146-1148   (version 4.3c)
//   SUBROUTINE TLIFT     climt/_lib/emanuel/convect43c.f90:1152-1219
// statement by statement with 1-based arrays, full (NL+1)^2 mixing matrices and the Fortran's loop bounds, and the
// column loop of the Cython shim (climt/_components/emanuel/_emanuel_convection.pyx:96-201).
// Restated for IPBL = 0 (the shim forces it, _emanuel_convection.pyx:71-73: the dry-adiabatic adjustment at
// convect43c.f90:336-423 never runs) and NTRA = 0 (climt/_components/emanuel/component.py:228).
//
// Pinning: the Fortran cannot be compiled here (no Fortran compiler) and the reference's own golden caches for this
// component hold only zeros (its default state does not convect).  The restatement is pinned instead against the
// reference's numba port of the same routine (climt/_components/emanuel/pure_python_v3.py:_convect_functional_np), RUN in
// the build container on convecting soundings -- tests/golden/make_emanuel_golden.py -> tests/golden/emanuel_reference.npz.
// The port differs from the Fortran in one constant: rain/snow fall-speed switch at T_freeze = 273.15 K
// (pure_python_v3.py:586-587) instead of 273.0 K (convect43c.f90:885); `t_rain` below selects it.
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "ftn.hpp"

namespace {

struct Par {
  double minorig, elcrit, tlcrit, entp, sigd, sigs, omtrain, omtsnow, coeffr, coeffs, cu, beta, dtmax, alpha, damp;
  double cpd, cpv, cl, rv, rd, lv0, g, rowl, delt0, t_rain;
};

using orc::A1;
using orc::A2;

// convect43c.f90:1152-1219
void tlift(const Par& c, const A1& P, const A1& T, const A1& Q, const A1& QS, const A1& GZ, int ICB, int NK, A1& TVP, A1& TPK, A1& CLW,
           int NL, int KK) {
  const double CPVMCL = c.cl - c.cpv;
  const double EPS = c.rd / c.rv;
  const double EPSI = 1. / EPS;
  const double AH0 = (c.cpd * (1. - Q(NK)) + c.cl * Q(NK)) * T(NK) + Q(NK) * (c.lv0 - CPVMCL * (T(NK) - 273.15)) + GZ(NK);
  const double CPP = c.cpd * (1. - Q(NK)) + Q(NK) * c.cpv;
  const double CPINV = 1. / CPP;
  if (KK == 1) {
    for (int I = 1; I <= ICB - 1; ++I) CLW(I) = 0.0;
    for (int I = NK; I <= ICB - 1; ++I) {
      TPK(I) = T(NK) - (GZ(I) - GZ(NK)) * CPINV;
      TVP(I) = TPK(I) * (1. + Q(NK) * EPSI);
    }
  }
  int NST = ICB, NSB = ICB;
  if (KK == 2) {
    NST = NL;
    NSB = ICB + 1;
  }
  for (int I = NSB; I <= NST; ++I) {
    double TG = T(I);
    double QG = QS(I);
    const double ALV = c.lv0 - CPVMCL * (T(I) - 273.15);
    for (int J = 1; J <= 2; ++J) {
      double S = c.cpd + ALV * ALV * QG / (c.rv * T(I) * T(I));
      S = 1. / S;
      const double AHG = c.cpd * TG + (c.cl - c.cpd) * Q(NK) * T(I) + ALV * QG + GZ(I);
      TG = TG + S * (AH0 - AHG);
      TG = std::max(TG, 35.0);
      const double TC = TG - 273.15;
      const double DENOM = 243.5 + TC;
      double ES;
      if (TC >= 0.0) ES = 6.112 * std::exp(17.67 * TC / DENOM);
      else ES = std::exp(23.33086 - 6111.72784 / TG + 0.15215 * std::log(TG));
      QG = EPS * ES / (P(I) - ES * (1. - EPS));
    }
    TPK(I) = (AH0 - (c.cl - c.cpd) * Q(NK) * T(I) - GZ(I) - ALV * QG) / c.cpd;
    CLW(I) = Q(NK) - QG;
    CLW(I) = std::max(0.0, CLW(I));
    const double RG = QG / (1. - Q(NK));
    TVP(I) = TPK(I) * (1. + RG * EPSI);
  }
}

struct ColOut {
  int iflag;
  double precip, wd, tprime, qprime, cbmf, cape;
};

// convect43c.f90:146-1148 for one column.  T..P have ND entries, PH ND+1; FT..FV ND entries (zeroed here).
void convect(const Par& c, const A1& T, const A1& Q, const A1& QS, const A1& U, const A1& V, const A1& P, const A1& PH, int ND, int NL,
             double DELT, A1& FT, A1& FQ, A1& FU, A1& FV, ColOut& o) {
  const int NA = ND + 2;
  const int MINORIG = (int)c.minorig;
  A1 M(NA), MP(NA), TVP(NA), TV(NA), WATER(NA), QP(NA), EP(NA), WT(NA), EVAP(NA), CLW(NA), SIGP(NA), TP(NA), CPN(NA), LV(NA),
      LVCP(NA), H(NA), HP(NA), GZ(NA), HM(NA), UP(NA), VP(NA);
  std::vector<int> NENT((size_t)NA + 1, 0);
  A2 UENT(NA, NA), VENT(NA, NA), QENT(NA, NA), ELIJ(NA, NA), MENT(NA, NA), SIJ(NA, NA);
  const double CPVMCL = c.cl - c.cpv;
  const double EPS = c.rd / c.rv;
  const double EPSI = 1. / EPS;
  const double GINV = 1.0 / c.g;
  const double DELTI = 1.0 / DELT;
  for (int I = 1; I <= ND; ++I) FT(I) = FQ(I) = FU(I) = FV(I) = 0.0;
  o.precip = o.wd = o.tprime = o.qprime = 0.0;
  o.cape = 0.0;  // OUTCAPE is intent(out) but unset on the early returns; the caller's array is zero-initialised (component.py:301-303)
  o.iflag = 0;
  double& CBMF = o.cbmf;
  // geopotential, heat capacity, static energies (:426-458)
  GZ(1) = 0.0;
  CPN(1) = c.cpd * (1. - Q(1)) + Q(1) * c.cpv;
  H(1) = T(1) * CPN(1);
  LV(1) = c.lv0 - CPVMCL * (T(1) - 273.15);
  HM(1) = LV(1) * Q(1);
  TV(1) = T(1) * (1. + Q(1) * EPSI - Q(1));
  double AHMIN = 1.0E12;
  int IHMIN = NL;
  for (int I = 2; I <= NL + 1; ++I) {
    const double TVX = T(I) * (1. + Q(I) * EPSI - Q(I));
    const double TVY = T(I - 1) * (1. + Q(I - 1) * EPSI - Q(I - 1));
    GZ(I) = GZ(I - 1) + 0.5 * c.rd * (TVX + TVY) * (P(I - 1) - P(I)) / PH(I);
    CPN(I) = c.cpd * (1. - Q(I)) + c.cpv * Q(I);
    H(I) = T(I) * CPN(I) + GZ(I);
    LV(I) = c.lv0 - CPVMCL * (T(I) - 273.15);
    HM(I) = (c.cpd * (1. - Q(I)) + c.cl * Q(I)) * (T(I) - T(1)) + LV(I) * Q(I) + GZ(I);
    TV(I) = T(I) * (1. + Q(I) * EPSI - Q(I));
    if (I >= MINORIG && HM(I) < AHMIN && HM(I) < HM(I - 1)) {
      AHMIN = HM(I);
      IHMIN = I;
    }
  }
  IHMIN = std::min(IHMIN, NL - 1);
  // level of maximum moist static energy below IHMIN (:463-470)
  double AHMAX = 0.0;
  int NK = 0;
  for (int I = MINORIG; I <= IHMIN; ++I)
    if (HM(I) > AHMAX) {
      NK = I;
      AHMAX = HM(I);
    }
  // The Fortran reads T(0) when no HM is positive (NK stays 0); the numba port then uses its level 0 = Fortran level 1.
  // Any sounding with HM(1) = LV*Q(1) > 0 sets NK >= 1; with Q(1) <= 0 follow the port (NK -> 1: Q(NK) <= 0 returns).
  if (NK == 0) NK = 1;
  if (T(NK) < 250.0 || Q(NK) <= 0.0 || IHMIN == (NL - 1)) {
    o.iflag = 0;
    CBMF = 0.0;
    return;
  }
  // lifted condensation level (:491-503)
  const double RH = Q(NK) / QS(NK);
  const double CHI = T(NK) / (1669.0 - 122.0 * RH - T(NK));
  const double PLCL = P(NK) * std::pow(RH, CHI);
  if (PLCL < 200.0 || PLCL >= 2000.0) {
    o.iflag = 2;
    CBMF = 0.0;
    return;
  }
  // first level above the LCL (:509-519)
  int ICB = NL - 1;
  for (int I = NK + 1; I <= NL; ++I)
    if (P(I) < PLCL) ICB = std::min(ICB, I);
  if (ICB >= (NL - 1)) {
    o.iflag = 3;
    CBMF = 0.0;
    return;
  }
  tlift(c, P, T, Q, QS, GZ, ICB, NK, TVP, TP, CLW, NL, 1);
  for (int I = NK; I <= ICB; ++I) TVP(I) = TVP(I) - TP(I) * Q(NK);
  if (CBMF == 0.0 && TVP(ICB) <= (TV(ICB) - c.dtmax)) {
    o.iflag = 0;
    return;
  }
  if (o.iflag != 4) o.iflag = 1;
  tlift(c, P, T, Q, QS, GZ, ICB, NK, TVP, TP, CLW, NL, 2);
  // precipitation efficiencies (:564-582)
  for (int I = 1; I <= NK; ++I) {
    EP(I) = 0.0;
    SIGP(I) = c.sigs;
  }
  for (int I = NK + 1; I <= NL; ++I) {
    const double TCA = TP(I) - 273.15;
    double ELACRIT;
    if (TCA >= 0.0) ELACRIT = c.elcrit;
    else ELACRIT = c.elcrit * (1.0 - TCA / c.tlcrit);
    ELACRIT = std::max(ELACRIT, 0.0);
    const double EPMAX = 0.999;
    EP(I) = EPMAX * (1.0 - ELACRIT / std::max(CLW(I), 1.0E-8));
    EP(I) = std::max(EP(I), 0.0);
    EP(I) = std::min(EP(I), EPMAX);
    SIGP(I) = c.sigs;
  }
  for (int I = ICB + 1; I <= NL; ++I) TVP(I) = TVP(I) - TP(I) * Q(NK);
  TVP(NL + 1) = TVP(NL) - (GZ(NL + 1) - GZ(NL)) / c.cpd;
  // initialisation of the work arrays (:594-626)
  for (int I = 1; I <= NL + 1; ++I) {
    HP(I) = H(I);
    NENT[I] = 0;
    WATER(I) = 0.0;
    EVAP(I) = 0.0;
    WT(I) = c.omtsnow;
    MP(I) = 0.0;
    M(I) = 0.0;
    LVCP(I) = LV(I) / CPN(I);
    for (int J = 1; J <= NL + 1; ++J) {
      QENT(I, J) = Q(J);
      ELIJ(I, J) = 0.0;
      MENT(I, J) = 0.0;
      SIJ(I, J) = 0.0;
      UENT(I, J) = U(J);
      VENT(I, J) = V(J);
    }
  }
  QP(1) = Q(1);
  UP(1) = U(1);
  VP(1) = V(1);
  for (int I = 2; I <= NL + 1; ++I) {
    QP(I) = Q(I - 1);
    UP(I) = U(I - 1);
    VP(I) = V(I - 1);
  }
  // level of neutral buoyancy and CAPE (:632-655)
  double CAPE = 0.0, CAPEM = 0.0;
  int INB = ICB + 1, INB1 = INB;
  double BYP = 0.0;
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
  INB = std::max(INB, INB1);
  CAPE = CAPEM + BYP;
  double DEFRAC = CAPEM - CAPE;
  DEFRAC = std::max(DEFRAC, 0.001);
  double FRAC = -CAPE / DEFRAC;
  FRAC = std::min(FRAC, 1.0);
  FRAC = std::max(FRAC, 0.0);
  o.cape = CAPE;
  for (int I = ICB; I <= INB; ++I) HP(I) = H(NK) + (LV(I) + (c.cpd - c.cpv) * T(I)) * EP(I) * CLW(I);
  // cloud base mass flux (:667-704)
  double DBOSUM = 0.0;
  const double TVPPLCL = TVP(ICB - 1) - c.rd * TVP(ICB - 1) * (P(ICB - 1) - PLCL) / (CPN(ICB - 1) * P(ICB - 1));
  const double TVAPLCL = TV(ICB) + (TVP(ICB) - TVP(ICB + 1)) * (PLCL - P(ICB)) / (P(ICB) - P(ICB + 1));
  double DTPBL = 0.0;
  for (int I = NK; I <= ICB - 1; ++I) DTPBL = DTPBL + (TVP(I) - TV(I)) * (PH(I) - PH(I + 1));
  DTPBL = DTPBL / (PH(NK) - PH(ICB));
  const double DTMIN = TVPPLCL - TVAPLCL + c.dtmax + DTPBL;
  const double DTMA = DTMIN;
  const double CBMFOLD = CBMF;
  const double DAMPS = c.damp * DELT / c.delt0;
  CBMF = (1. - DAMPS) * CBMF + 0.1 * c.alpha * DTMA;
  CBMF = std::max(CBMF, 0.0);
  if (CBMF == 0.0 && CBMFOLD == 0.0) return;
  // rates of mixing (:708-718)
  M(ICB) = 0.0;
  for (int I = ICB + 1; I <= INB; ++I) {
    const int K = std::min(I, INB1);
    const double DBO = std::fabs(TV(K) - TVP(K)) + c.entp * 0.02 * (PH(K) - PH(K + 1));
    DBOSUM = DBOSUM + DBO;
    M(I) = CBMF * DBO;
  }
  for (int I = ICB + 1; I <= INB; ++I) M(I) = M(I) / DBOSUM;
  // entrained air mass flux, total water, condensed water, mixing fraction (:727-786)
  for (int I = ICB + 1; I <= INB; ++I) {
    const double QTI = Q(NK) - EP(I) * CLW(I);
    for (int J = ICB; J <= INB; ++J) {
      const double BF2 = 1. + LV(J) * LV(J) * QS(J) / (c.rv * T(J) * T(J) * c.cpd);
      double ANUM = H(J) - HP(I) + (c.cpv - c.cpd) * T(J) * (QTI - Q(J));
      double DENOM = H(I) - HP(I) + (c.cpd - c.cpv) * (Q(I) - QTI) * T(J);
      double DEI = DENOM;
      if (std::fabs(DEI) < 0.01) DEI = 0.01;
      SIJ(I, J) = ANUM / DEI;
      SIJ(I, I) = 1.0;
      double ALTEM = SIJ(I, J) * Q(I) + (1. - SIJ(I, J)) * QTI - QS(J);
      ALTEM = ALTEM / BF2;
      const double CWAT = CLW(J) * (1. - EP(J));
      const double STEMP = SIJ(I, J);
      if ((STEMP < 0.0 || STEMP > 1.0 || ALTEM > CWAT) && J > I) {
        ANUM = ANUM - LV(J) * (QTI - QS(J) - CWAT * BF2);
        DENOM = DENOM + LV(J) * (Q(I) - QTI);
        if (std::fabs(DENOM) < 0.01) DENOM = 0.01;
        SIJ(I, J) = ANUM / DENOM;
        ALTEM = SIJ(I, J) * Q(I) + (1. - SIJ(I, J)) * QTI - QS(J);
        ALTEM = ALTEM - (BF2 - 1.) * CWAT;
      }
      if (SIJ(I, J) > 0.0 && SIJ(I, J) < 0.9) {
        QENT(I, J) = SIJ(I, J) * Q(I) + (1. - SIJ(I, J)) * QTI;
        UENT(I, J) = SIJ(I, J) * U(I) + (1. - SIJ(I, J)) * U(NK);
        VENT(I, J) = SIJ(I, J) * V(I) + (1. - SIJ(I, J)) * V(NK);
        ELIJ(I, J) = ALTEM;
        ELIJ(I, J) = std::max(0.0, ELIJ(I, J));
        MENT(I, J) = M(I) / (1. - SIJ(I, J));
        NENT[I] = NENT[I] + 1;
      }
      SIJ(I, J) = std::max(0.0, SIJ(I, J));
      SIJ(I, J) = std::min(1.0, SIJ(I, J));
    }
    if (NENT[I] == 0) {
      MENT(I, I) = M(I);
      QENT(I, I) = Q(NK) - EP(I) * CLW(I);
      UENT(I, I) = U(NK);
      VENT(I, I) = V(NK);
      ELIJ(I, I) = CLW(I);
      SIJ(I, I) = 1.0;
    }
  }
  SIJ(INB, INB) = 1.0;
  // normalise the entrained fluxes to equal probabilities of mixing (:792-856)
  for (int I = ICB + 1; I <= INB; ++I) {
    if (NENT[I] != 0) {
      const double QP1 = Q(NK) - EP(I) * CLW(I);
      const double ANUM = H(I) - HP(I) - LV(I) * (QP1 - QS(I));
      double DENOM = H(I) - HP(I) + LV(I) * (Q(I) - QP1);
      if (std::fabs(DENOM) < 0.01) DENOM = 0.01;
      double SCRIT = ANUM / DENOM;
      const double ALT = QP1 - QS(I) + SCRIT * (Q(I) - QP1);
      if (ALT < 0.0) SCRIT = 1.0;
      SCRIT = std::max(SCRIT, 0.0);
      double ASIJ = 0.0;
      double SMIN = 1.0;
      for (int J = ICB; J <= INB; ++J) {
        if (SIJ(I, J) > 0.0 && SIJ(I, J) < 0.9) {
          double SMID, SJMAX, SJMIN;
          if (J > I) {
            SMID = std::min(SIJ(I, J), SCRIT);
            SJMAX = SMID;
            SJMIN = SMID;
            if (SMID < SMIN && SIJ(I, J + 1) < SMID) {
              SMIN = SMID;
              SJMAX = std::min(std::min(SIJ(I, J + 1), SIJ(I, J)), SCRIT);
              SJMIN = std::max(SIJ(I, J - 1), SIJ(I, J));
              SJMIN = std::min(SJMIN, SCRIT);
            }
          } else {
            SJMAX = std::max(SIJ(I, J + 1), SCRIT);
            SMID = std::max(SIJ(I, J), SCRIT);
            SJMIN = 0.0;
            if (J > 1) SJMIN = SIJ(I, J - 1);
            SJMIN = std::max(SJMIN, SCRIT);
          }
          const double DELP = std::fabs(SJMAX - SMID);
          const double DELM = std::fabs(SJMIN - SMID);
          ASIJ = ASIJ + (DELP + DELM) * (PH(J) - PH(J + 1));
          MENT(I, J) = MENT(I, J) * (DELP + DELM) * (PH(J) - PH(J + 1));
        }
      }
      ASIJ = std::max(1.0E-21, ASIJ);
      ASIJ = 1.0 / ASIJ;
      for (int J = ICB; J <= INB; ++J) MENT(I, J) = MENT(I, J) * ASIJ;
      double BSUM = 0.0;
      for (int J = ICB; J <= INB; ++J) BSUM = BSUM + MENT(I, J);
      if (BSUM < 1.0E-18) {
        NENT[I] = 0;
        MENT(I, I) = M(I);
        QENT(I, I) = Q(NK) - EP(I) * CLW(I);
        UENT(I, I) = U(NK);
        VENT(I, I) = V(NK);
        ELIJ(I, I) = CLW(I);
        SIJ(I, I) = 1.0;
      }
    }
  }
  // precipitating downdraft (:866-956)
  if (!(EP(INB) < 0.0001)) {
    int JTT = 2;
    for (int I = INB; I >= 1; --I) {
      double WDTRAIN = c.g * EP(I) * M(I) * CLW(I);
      if (I > 1) {
        for (int J = 1; J <= I - 1; ++J) {
          double AWAT = ELIJ(J, I) - (1. - EP(I)) * CLW(I);
          AWAT = std::max(0.0, AWAT);
          WDTRAIN = WDTRAIN + c.g * AWAT * MENT(J, I);
        }
      }
      double COEFF = c.coeffs;
      WT(I) = c.omtsnow;
      if (T(I) > c.t_rain) {
        COEFF = c.coeffr;
        WT(I) = c.omtrain;
      }
      const double QSM = 0.5 * (Q(I) + QP(I + 1));
      double AFAC = COEFF * PH(I) * (QS(I) - QSM) / (1.0E4 + 2.0E3 * PH(I) * QS(I));
      AFAC = std::max(AFAC, 0.0);
      double SIGT = SIGP(I);
      SIGT = std::max(0.0, SIGT);
      SIGT = std::min(1.0, SIGT);
      const double B6 = 100. * (PH(I) - PH(I + 1)) * SIGT * AFAC / WT(I);
      const double C6 = (WATER(I + 1) * WT(I + 1) + WDTRAIN / c.sigd) / WT(I);
      const double REVAP = 0.5 * (-B6 + std::sqrt(B6 * B6 + 4. * C6));
      EVAP(I) = SIGT * AFAC * REVAP;
      WATER(I) = REVAP * REVAP;
      if (I != 1) {
        double DHDP = (H(I) - H(I - 1)) / (P(I - 1) - P(I));
        DHDP = std::max(DHDP, 10.0);
        MP(I) = 100. * GINV * LV(I) * c.sigd * EVAP(I) / DHDP;
        MP(I) = std::max(MP(I), 0.0);
        const double FAC = 20.0 / (PH(I - 1) - PH(I));
        MP(I) = (FAC * MP(I + 1) + MP(I)) / (1. + FAC);
        if (P(I) > (0.949 * P(1))) {
          JTT = std::max(JTT, I);
          MP(I) = MP(JTT) * (P(1) - P(I)) / (P(1) - P(JTT));
        }
      }
      if (I == INB) continue;
      double QSTM;
      if (I == 1) QSTM = QS(1);
      else QSTM = QS(I - 1);
      if (MP(I) > MP(I + 1)) {
        const double RAT = MP(I + 1) / MP(I);
        QP(I) = QP(I + 1) * RAT + Q(I) * (1.0 - RAT) + 100. * GINV * c.sigd * (PH(I) - PH(I + 1)) * (EVAP(I) / MP(I));
        UP(I) = UP(I + 1) * RAT + U(I) * (1. - RAT);
        VP(I) = VP(I + 1) * RAT + V(I) * (1. - RAT);
      } else {
        if (MP(I + 1) > 0.0) {
          QP(I) = (GZ(I + 1) - GZ(I) + QP(I + 1) * (LV(I + 1) + T(I + 1) * (c.cl - c.cpd)) + c.cpd * (T(I + 1) - T(I))) /
                  (LV(I) + T(I) * (c.cl - c.cpd));
          UP(I) = UP(I + 1);
          VP(I) = VP(I + 1);
        }
      }
      QP(I) = std::min(QP(I), QSTM);
      QP(I) = std::max(QP(I), 0.0);
    }
    o.precip = o.precip + WT(1) * c.sigd * WATER(1) * 3600. * 24000. / (c.rowl * c.g);
  }
  // downdraft velocity scale, surface fluctuations (:966-968)
  o.wd = c.beta * std::fabs(MP(ICB)) * 0.01 * c.rd * T(ICB) / (c.sigd * P(ICB));
  o.qprime = 0.5 * (QP(1) - Q(1));
  o.tprime = c.lv0 * o.qprime / c.cpd;
  // tendencies of the lowest level (:974-1003)
  double DPINV = 0.01 / (PH(1) - PH(2));
  double AM = 0.0;
  if (NK == 1)
    for (int K = 2; K <= INB; ++K) AM = AM + M(K);
  if ((2. * c.g * DPINV * AM) >= DELTI) o.iflag = 4;
  FT(1) = FT(1) + c.g * DPINV * AM * (T(2) - T(1) + (GZ(2) - GZ(1)) / CPN(1));
  FT(1) = FT(1) - LVCP(1) * c.sigd * EVAP(1);
  FT(1) = FT(1) + c.sigd * WT(2) * (c.cl - c.cpd) * WATER(2) * (T(2) - T(1)) * DPINV / CPN(1);
  FQ(1) = FQ(1) + c.g * MP(2) * (QP(2) - Q(1)) * DPINV + c.sigd * EVAP(1);
  FQ(1) = FQ(1) + c.g * AM * (Q(2) - Q(1)) * DPINV;
  FU(1) = FU(1) + c.g * DPINV * (MP(2) * (UP(2) - U(1)) + AM * (U(2) - U(1)));
  FV(1) = FV(1) + c.g * DPINV * (MP(2) * (VP(2) - V(1)) + AM * (V(2) - V(1)));
  for (int J = 2; J <= INB; ++J) {
    FQ(1) = FQ(1) + c.g * DPINV * MENT(J, 1) * (QENT(J, 1) - Q(1));
    FU(1) = FU(1) + c.g * DPINV * MENT(J, 1) * (UENT(J, 1) - U(1));
    FV(1) = FV(1) + c.g * DPINV * MENT(J, 1) * (VENT(J, 1) - V(1));
  }
  // tendencies above the lowest level (:1012-1086)
  for (int I = 2; I <= INB; ++I) {
    DPINV = 0.01 / (PH(I) - PH(I + 1));
    const double CPINV = 1.0 / CPN(I);
    double AMP1 = 0.0;
    double AD = 0.0;
    if (I >= NK)
      for (int K = I + 1; K <= INB + 1; ++K) AMP1 = AMP1 + M(K);
    for (int K = 1; K <= I; ++K)
      for (int J = I + 1; J <= INB + 1; ++J) AMP1 = AMP1 + MENT(K, J);
    if ((2. * c.g * DPINV * AMP1) >= DELTI) o.iflag = 4;
    for (int K = 1; K <= I - 1; ++K)
      for (int J = I; J <= INB; ++J) AD = AD + MENT(J, K);
    FT(I) = FT(I) + c.g * DPINV * (AMP1 * (T(I + 1) - T(I) + (GZ(I + 1) - GZ(I)) * CPINV) - AD * (T(I) - T(I - 1) + (GZ(I) - GZ(I - 1)) * CPINV)) -
            c.sigd * LVCP(I) * EVAP(I);
    FT(I) = FT(I) + c.g * DPINV * MENT(I, I) * (HP(I) - H(I) + T(I) * (c.cpv - c.cpd) * (Q(I) - QENT(I, I))) * CPINV;
    FT(I) = FT(I) + c.sigd * WT(I + 1) * (c.cl - c.cpd) * WATER(I + 1) * (T(I + 1) - T(I)) * DPINV * CPINV;
    FQ(I) = FQ(I) + c.g * DPINV * (AMP1 * (Q(I + 1) - Q(I)) - AD * (Q(I) - Q(I - 1)));
    FU(I) = FU(I) + c.g * DPINV * (AMP1 * (U(I + 1) - U(I)) - AD * (U(I) - U(I - 1)));
    FV(I) = FV(I) + c.g * DPINV * (AMP1 * (V(I + 1) - V(I)) - AD * (V(I) - V(I - 1)));
    for (int K = 1; K <= I - 1; ++K) {
      double AWAT = ELIJ(K, I) - (1. - EP(I)) * CLW(I);
      AWAT = std::max(AWAT, 0.0);
      FQ(I) = FQ(I) + c.g * DPINV * MENT(K, I) * (QENT(K, I) - AWAT - Q(I));
      FU(I) = FU(I) + c.g * DPINV * MENT(K, I) * (UENT(K, I) - U(I));
      FV(I) = FV(I) + c.g * DPINV * MENT(K, I) * (VENT(K, I) - V(I));
    }
    for (int K = I; K <= INB; ++K) {
      FQ(I) = FQ(I) + c.g * DPINV * MENT(K, I) * (QENT(K, I) - Q(I));
      FU(I) = FU(I) + c.g * DPINV * MENT(K, I) * (UENT(K, I) - U(I));
      FV(I) = FV(I) + c.g * DPINV * MENT(K, I) * (VENT(K, I) - V(I));
    }
    FQ(I) = FQ(I) + c.sigd * EVAP(I) + c.g * (MP(I + 1) * (QP(I + 1) - Q(I)) - MP(I) * (QP(I) - Q(I - 1))) * DPINV;
    FU(I) = FU(I) + c.g * (MP(I + 1) * (UP(I + 1) - U(I)) - MP(I) * (UP(I) - U(I - 1))) * DPINV;
    FV(I) = FV(I) + c.g * (MP(I + 1) * (VP(I + 1) - V(I)) - MP(I) * (VP(I) - V(I - 1))) * DPINV;
  }
  // top-of-convection adjustment to the level of zero CAPE (:1092-1116)
  const double FQOLD = FQ(INB);
  FQ(INB) = FQ(INB) * (1. - FRAC);
  FQ(INB - 1) = FQ(INB - 1) + FRAC * FQOLD * ((PH(INB) - PH(INB + 1)) / (PH(INB - 1) - PH(INB))) * LV(INB) / LV(INB - 1);
  const double FTOLD = FT(INB);
  FT(INB) = FT(INB) * (1. - FRAC);
  FT(INB - 1) = FT(INB - 1) + FRAC * FTOLD * ((PH(INB) - PH(INB + 1)) / (PH(INB - 1) - PH(INB))) * CPN(INB) / CPN(INB - 1);
  const double FUOLD = FU(INB);
  FU(INB) = FU(INB) * (1. - FRAC);
  FU(INB - 1) = FU(INB - 1) + FRAC * FUOLD * ((PH(INB) - PH(INB + 1)) / (PH(INB - 1) - PH(INB)));
  const double FVOLD = FV(INB);
  FV(INB) = FV(INB) * (1. - FRAC);
  FV(INB - 1) = FV(INB - 1) + FRAC * FVOLD * ((PH(INB) - PH(INB + 1)) / (PH(INB - 1) - PH(INB)));
  // exact enthalpy and momentum conservation (:1122-1136)
  double ENTS = 0.0, UAV = 0.0, VAV = 0.0;
  for (int I = 1; I <= INB; ++I) {
    ENTS = ENTS + (CPN(I) * FT(I) + LV(I) * FQ(I)) * (PH(I) - PH(I + 1));
    UAV = UAV + FU(I) * (PH(I) - PH(I + 1));
    VAV = VAV + FV(I) * (PH(I) - PH(I + 1));
  }
  ENTS = ENTS / (PH(1) - PH(INB + 1));
  UAV = UAV / (PH(1) - PH(INB + 1));
  VAV = VAV / (PH(1) - PH(INB + 1));
  for (int I = 1; I <= INB; ++I) {
    FT(I) = FT(I) - ENTS / CPN(I);
    FU(I) = (1. - c.cu) * (FU(I) - UAV);
    FV(I) = (1. - c.cu) * (FV(I) - VAV);
  }
}

}  // namespace

extern "C" {

// The column loop of _emanuel_convection.pyx:convect (:148-201).  Arrays are the component's: (ncol, nlev) C order, level 1
// at the surface, pressures in mbar; cbmf is read and written; iflag is int32.  par: the 25 doubles of `Par` in order.
int orc_emanuel_convect(const double* par, int ncol, int nlev, int max_conv_lev, double dt, const double* t, const double* q,
                        const double* qs, const double* u, const double* v, const double* p, const double* ph, double* cbmf,
                        int32_t* iflag, double* ft, double* fq, double* fu, double* fv, double* precip, double* wd, double* tprime,
                        double* qprime, double* cape) {
  Par c;
  static_assert(sizeof(Par) == 25 * sizeof(double), "Par layout");
  std::memcpy(&c, par, sizeof(Par));
  const int ND = nlev, NL = max_conv_lev;
  A1 T(ND + 2), Q(ND + 2), QS(ND + 2), U(ND + 2), V(ND + 2), P(ND + 2), PH(ND + 3), FT(ND + 2), FQ(ND + 2), FU(ND + 2), FV(ND + 2);
  for (int col = 0; col < ncol; ++col) {
    const size_t o = (size_t)col * nlev, oi = (size_t)col * (nlev + 1);
    for (int k = 0; k < nlev; ++k) {
      T(k + 1) = t[o + k]; Q(k + 1) = q[o + k]; QS(k + 1) = qs[o + k]; U(k + 1) = u[o + k]; V(k + 1) = v[o + k]; P(k + 1) = p[o + k];
    }
    for (int k = 0; k <= nlev; ++k) PH(k + 1) = ph[oi + k];
    ColOut out{};
    out.cbmf = cbmf[col];
    convect(c, T, Q, QS, U, V, P, PH, ND, NL, dt, FT, FQ, FU, FV, out);
    for (int k = 0; k < nlev; ++k) {
      ft[o + k] = FT(k + 1); fq[o + k] = FQ(k + 1); fu[o + k] = FU(k + 1); fv[o + k] = FV(k + 1);
    }
    cbmf[col] = out.cbmf; iflag[col] = out.iflag; precip[col] = out.precip; wd[col] = out.wd; tprime[col] = out.tprime;
    qprime[col] = out.qprime; cape[col] = out.cape;
  }
  return 0;
}

}  // extern "C"
