// fastfmt.h -- allocation-free integer / "%f" text generation for the freq and view writers
// (a 30x human `freq` prints ~56 M rows; fprintf("%f") dominates the host side).
//
// fmt_f6() produces exactly what printf("%f", x) prints for any finite x in [0, 2^63): the double is taken
// apart into mantissa * 2^exp, multiplied by 10^6 in 128-bit integers and rounded to nearest, ties to even,
// on its EXACT value -- which is what glibc's printf does.  tests/test_fastfmt.py checks it against snprintf
// exhaustively for every n_mod/n_called and n_mod*100/n_called with n_called <= 1024 and on millions of
// random 32-bit pairs.
#ifndef MMH_FASTFMT_H
#define MMH_FASTFMT_H
#include <stdint.h>
#include <string.h>

namespace mmh {

inline char *fmt_u64(char *p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
inline char *fmt_i32(char *p, int32_t v) {
    if (v < 0) { *p++ = '-'; return fmt_u64(p, (uint64_t)(-(int64_t)v)); }
    return fmt_u64(p, (uint64_t)v);
}
inline char *fmt_str(char *p, const char *s) { size_t n = strlen(s); memcpy(p, s, n); return p + n; }

// == sprintf(p, "%f", x) for finite 0 <= x < 2^63; returns the end of the text (no NUL)
inline char *fmt_f6(char *p, double x) {
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const int e = (int)((bits >> 52) & 0x7ff);
    uint64_t mant = bits & ((1ull << 52) - 1);
    int exp2;
    if (e == 0) exp2 = -1074; else { mant |= 1ull << 52; exp2 = e - 1075; }      // x == mant * 2^exp2
    uint64_t ip, fp;                                                              // integer part, 6 fraction digits
    if (exp2 >= 0) { ip = mant << exp2; fp = 0; }
    else {
        const int sh = -exp2;
        unsigned __int128 t = (unsigned __int128)mant * 1000000u, q;              // < 2^73
        if (sh >= 100) q = 0;                                                     // t / 2^sh < 2^-27: rounds to 0
        else {
            q = t >> sh;
            const unsigned __int128 rem = t & (((unsigned __int128)1 << sh) - 1), half = (unsigned __int128)1 << (sh - 1);
            if (rem > half || (rem == half && (q & 1))) ++q;                      // to nearest, ties to even, on the exact value
        }
        ip = (uint64_t)(q / 1000000u); fp = (uint64_t)(q % 1000000u);
    }
    p = fmt_u64(p, ip);
    *p++ = '.';
    for (int k = 5; k >= 0; --k) { p[k] = (char)('0' + fp % 10); fp /= 10; }
    return p + 6;
}

}  // namespace mmh
#endif
