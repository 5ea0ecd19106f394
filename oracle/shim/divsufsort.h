/*
 * TEST INFRASTRUCTURE ONLY (oracle/): declarations of the two parallel-divsufsort
 * entry points the reference builder calls (gsa.cpp:22,33).  The library itself is not
 * vendored by the reference and is absent here; oracle/sa_standin.cpp supplies a
 * from-scratch SA-IS implementation with the same signatures.  Build-side only; the
 * query path never touches it.
 */
#ifndef ORACLE_SHIM_DIVSUFSORT_H
#define ORACLE_SHIM_DIVSUFSORT_H
#include <cstdint>
int divsufsort(const uint8_t *T, int64_t *SA, int64_t n);
int sufcheck(const uint8_t *T, const int64_t *SA, int64_t n, bool verbose);
#endif
