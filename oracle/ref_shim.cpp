// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors onto the REFERENCE's own Davies/Liu code, compiled in place from
// /root/reference/regression/{MixtureChiSquare.cpp (which #includes qfc.c), cdflib.cpp}
// into oracle/_ref/libmixchisq_ref.so by oracle/Makefile.  Nothing from the reference is
// copied into this repository; this file only calls its public class
// (regression/MixtureChiSquare.h:7-76).
#include "MixtureChiSquare.h"

// qf() is defined (C++ linkage) by the reference translation unit MixtureChiSquare.cpp, which
// #includes qfc.c (regression/MixtureChiSquare.cpp:5, regression/qfc.c:304).
double qf(double*, double*, int*, int, double, double, int, double, double*, int*);

extern "C" {
// MixtureChiSquare::getPvalue (regression/MixtureChiSquare.cpp:7-29): -1 on Davies fault.
double ref_mixchisq_pvalue(const double* lambda, int n, double Q) {
  MixtureChiSquare m;
  for (int i = 0; i < n; ++i) m.addLambda(lambda[i]);
  return m.getPvalue(Q);
}
// MixtureChiSquare::getLiuPvalue (regression/MixtureChiSquare.cpp:44-83)
double ref_liu_pvalue(const double* lambda, int n, double Q) {
  MixtureChiSquare m;
  for (int i = 0; i < n; ++i) m.addLambda(lambda[i]);
  return m.getLiuPvalue(Q);
}
// raw qf() (regression/qfc.c:304) with its fault code and trace[]
double ref_qf(double* lb, double* nc, int* n, int r, double sigma, double c, int lim, double acc,
              double* trace, int* ifault) {
  return qf(lb, nc, n, r, sigma, c, lim, acc, trace, ifault);
}
}
