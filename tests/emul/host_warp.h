// TEST INFRASTRUCTURE: 32-wide vector backend that lets g++ run the warp-cooperative kernel
// bodies of edlib_b200/csrc/eb_core.h on the CPU (the device backend in eb_kernels.cu maps the
// same names to shuffles, votes and plain registers).  Used only by tests/emul/emul_backend.cpp.
#pragma once
#include <stdint.h>
#include <type_traits>

#include "eb_common.h"

namespace ebhost {

struct PV {
    bool x[32];
};
inline PV operator&(const PV& a, const PV& b) { PV r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] && b.x[i]; return r; }
inline PV operator|(const PV& a, const PV& b) { PV r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] || b.x[i]; return r; }
inline PV operator!(const PV& a) { PV r; for (int i = 0; i < 32; ++i) r.x[i] = !a.x[i]; return r; }

struct V {
    uint32_t x[32];
    V() {}
    template <class T, class = typename std::enable_if<std::is_integral<T>::value>::type>
    explicit V(T s) { for (int i = 0; i < 32; ++i) x[i] = (uint32_t)s; }
};
#define EBH_BIN(op)                                                                                         \
    inline V operator op(const V& a, const V& b) { V r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] op b.x[i]; return r; }
EBH_BIN(&) EBH_BIN(|) EBH_BIN(^) EBH_BIN(+) EBH_BIN(-) EBH_BIN(*)
#undef EBH_BIN
inline V operator~(const V& a) { V r; for (int i = 0; i < 32; ++i) r.x[i] = ~a.x[i]; return r; }
inline V operator<<(const V& a, int s) { V r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] << s; return r; }
inline V operator>>(const V& a, int s) { V r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] >> s; return r; }
inline V operator>>(const V& a, const V& s) { V r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] >> (s.x[i] & 31); return r; }
#define EBH_CMP(op)                                                                                          \
    inline PV operator op(const V& a, const V& b) { PV r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] op b.x[i]; return r; } \
    inline PV operator op(const V& a, uint32_t b) { PV r; for (int i = 0; i < 32; ++i) r.x[i] = a.x[i] op b; return r; }
EBH_CMP(==) EBH_CMP(!=) EBH_CMP(<) EBH_CMP(<=)
#undef EBH_CMP

struct HostWarp {
    using U = V;
    using P = PV;
    static U lane() { V r; for (int i = 0; i < 32; ++i) r.x[i] = (uint32_t)i; return r; }
    template <class F> static U map(const U& a, F f) { V r; for (int i = 0; i < 32; ++i) r.x[i] = f(a.x[i]); return r; }
    static U sel(const P& p, const U& a, const U& b) { V r; for (int i = 0; i < 32; ++i) r.x[i] = p.x[i] ? a.x[i] : b.x[i]; return r; }
    static U toU(const P& p) { V r; for (int i = 0; i < 32; ++i) r.x[i] = p.x[i] ? 1u : 0u; return r; }
    static uint32_t ballot(const P& p) { uint32_t m = 0; for (int i = 0; i < 32; ++i) m |= (p.x[i] ? 1u : 0u) << i; return m; }
    static bool any(const P& p) { return ballot(p) != 0; }
    static U shfl_up1(const U& a) { V r; r.x[0] = a.x[0]; for (int i = 1; i < 32; ++i) r.x[i] = a.x[i - 1]; return r; }
    static U shfl_down1(const U& a) { V r; r.x[31] = a.x[31]; for (int i = 0; i < 31; ++i) r.x[i] = a.x[i + 1]; return r; }
    static uint32_t bcast(const U& a, int srcLane) { return a.x[srcLane & 31]; }
    static U gather8(const uint8_t* base, const U& idx, const P& ok) { V r; for (int i = 0; i < 32; ++i) r.x[i] = ok.x[i] ? base[idx.x[i]] : 0u; return r; }
    static U gather8_neg(const uint8_t* base, const U& idx, const P& ok) { V r; for (int i = 0; i < 32; ++i) r.x[i] = ok.x[i] ? *(base - (ptrdiff_t)idx.x[i]) : 0u; return r; }
    static U gather32(const uint32_t* base, const U& idx, const P& ok) { V r; for (int i = 0; i < 32; ++i) r.x[i] = ok.x[i] ? base[idx.x[i]] : 0u; return r; }
    static void scatterU2(eb::U2* base, const U& idx, const U& a, const U& b, const P& ok) {
        for (int i = 0; i < 32; ++i) if (ok.x[i]) { base[idx.x[i]].x = a.x[i]; base[idx.x[i]].y = b.x[i]; }
    }
    static void scatter8(uint8_t* base, const U& idx, const U& v, const P& ok) {
        for (int i = 0; i < 32; ++i) if (ok.x[i]) base[idx.x[i]] = (uint8_t)v.x[i];
    }
    static void store_uniform(int* p, int v) { *p = v; }
    static int atomic_add_uniform(int* p, int v) { int o = *p; *p = o + v; return o; }
    template <int R>
    static void dump_column(int* out, const U (&Pv)[R], const U (&Mv)[R], const U& sb, int topChunk, int, int off, int m);
};

}  // namespace ebhost
