/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * CPU restatement ("oracle") of the reference algorithm behind edlibAlign()
 * (reference: edlib/src/edlib.cpp, commit 0ddc23e / v1.2.6).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product library (edlib_b200/lib/libedlib_b200.so) never links, loads or calls it.
 *
 * Parity pinning: the oracle is checked field-for-field against the UNMODIFIED reference
 * compiled from /root/reference into oracle/_ref/libedlib_ref.so (tests/test_oracle.py),
 * against the reference's own hand vectors (runTests.cpp test1-16, cigar, equality and
 * empty-sequence cases; bindings/python/test.py known answers) and against the committed
 * golden fixtures in tests/golden/ that were generated from the reference.
 */
#ifndef EDLIB_ORACLE_H
#define EDLIB_ORACLE_H

#include "../include/edlib.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same contract as edlibAlign (ref edlib.cpp:146-301). */
EdlibAlignResult oracleAlign(const char* query, int queryLength,
                             const char* target, int targetLength,
                             const EdlibAlignConfig config);

/* Same contract as edlibAlignmentToCigar (ref edlib.cpp:303-350). */
char* oracleAlignmentToCigar(const unsigned char* alignment, int alignmentLength,
                             EdlibCigarFormat cigarFormat);

/* Same contract as edlibFreeAlignResult (ref edlib.cpp:1481-1485). */
void oracleFreeAlignResult(EdlibAlignResult result);

#ifdef __cplusplus
}
#endif
#endif
