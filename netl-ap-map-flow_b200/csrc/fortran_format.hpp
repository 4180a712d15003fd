// fortran_format.hpp -- gfortran-compatible numeric edit descriptors for the drop-in output files.
//
// The reference writes every file with Fortran formatted I/O (G15.8 for CSV maps,
// apmapflow/src/APERTURE_MAP_FLOW.F:905-915; G20.12 for the stat files :1184,1207; E14.6 for VTK
// :1245, OUTPUT_MODULE.F:328; F/E/G/I in the log).  The golden test diffs those files line by line
// (test/fortran/bin/test_model_code_coverage:115-155), so the digits, padding, "-0.0000000" and the
// three-digit-exponent form ("0.1027-100") have to match.  Digits come from printf's correctly
// rounded conversion, which is also what libgfortran uses internally.
#pragma once

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace ffmt {

inline std::string stars(int w) { return std::string((size_t)(w > 0 ? w : 1), '*'); }

inline std::string rjust(const std::string &s, int w)
{
    if ((int)s.size() >= w) return s;
    return std::string((size_t)(w - (int)s.size()), ' ') + s;
}

// Fw.d
inline std::string F(int w, int d, double x)
{
    if (std::isnan(x)) return rjust("NaN", w);
    if (std::isinf(x)) return rjust(x > 0 ? "Infinity" : "-Infinity", w).size() > (size_t)w ? rjust(x > 0 ? "Inf" : "-Inf", w) : rjust(x > 0 ? "Infinity" : "-Infinity", w);
    char buf[512];
    snprintf(buf, sizeof buf, "%.*f", d, x);
    std::string s(buf);
    if (d == 0) s += '.';      // Fortran F editing always shows the decimal point ("12345678." for F11.0), printf does not
    if ((int)s.size() > w) {   // gfortran drops the optional leading zero when the field is too narrow
        if (s.compare(0, 2, "0.") == 0) s = s.substr(1);
        else if (s.compare(0, 3, "-0.") == 0) s = "-" + s.substr(2);
    }
    if ((int)s.size() > w) return stars(w);
    return rjust(s, w);
}

// Ew.d (scale factor 0): [-]0.ddddE+ee, exponent "+eee" without the E when it needs three digits
inline std::string E(int w, int d, double x)
{
    if (std::isnan(x)) return rjust("NaN", w);
    if (std::isinf(x)) return rjust(x > 0 ? "Infinity" : "-Infinity", w);
    char buf[512];
    const bool neg = std::signbit(x);
    const double a = std::fabs(x);
    std::string digits;
    int ex = 0;
    if (a == 0.0) {
        digits.assign((size_t)d, '0');
        ex = 0;
    } else {
        snprintf(buf, sizeof buf, "%.*e", d - 1, a);   // D.DDDDe+XX
        const char *e = strchr(buf, 'e');
        ex = atoi(e + 1) + 1;
        for (const char *p = buf; p < e; p++)
            if (*p != '.') digits.push_back(*p);
    }
    char ebuf[16];
    if (std::abs(ex) < 100) snprintf(ebuf, sizeof ebuf, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
    else snprintf(ebuf, sizeof ebuf, "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
    std::string s = std::string(neg ? "-" : "") + "0." + digits + ebuf;
    if ((int)s.size() > w) s = std::string(neg ? "-" : "") + "." + digits + ebuf;
    if ((int)s.size() > w) return stars(w);
    return rjust(s, w);
}

// Gw.d: F(w-4).(d-s) followed by four blanks when the value rounded to d digits has
// 0 <= s <= d digits before the point (0.1 <= |x| < 10^d), otherwise Ew.d
inline std::string G(int w, int d, double x)
{
    if (std::isnan(x) || std::isinf(x)) return E(w, d, x);
    const double a = std::fabs(x);
    if (a == 0.0) return F(w - 4, d - 1, x) + "    ";
    char buf[64];
    snprintf(buf, sizeof buf, "%.*e", d - 1, a);
    const int s = atoi(strchr(buf, 'e') + 1) + 1;
    if (s < 0 || s > d) return E(w, d, x);
    return F(w - 4, d - s, x) + "    ";
}

// Iw.m
inline std::string I(int w, int m, long long v)
{
    char buf[96];
    snprintf(buf, sizeof buf, "%0*lld", m > 0 ? (m < 40 ? m : 40) : 1, v < 0 ? -v : v);
    std::string s = std::string(v < 0 ? "-" : "") + buf;
    if (w == 0) return s;
    if ((int)s.size() > w) return stars(w);
    return rjust(s, w);
}

// Aw for output: right justified in w columns (truncated on the right if longer)
inline std::string A(int w, const std::string &s)
{
    if ((int)s.size() >= w) return s.substr(0, (size_t)w);
    return rjust(s, w);
}

// ---- the same edit descriptors for the bulk writers: std::to_chars (correctly rounded like printf,
// several times faster) straight into the row buffer, no temporary strings.  tests/test_abi.py checks
// them against F/E/G above on random and edge-case values.
inline void append_padded(std::string &out, int w, const char *s, int n)
{
    if (n < w) out.append((size_t)(w - n), ' ');
    out.append(s, (size_t)n);
}

inline void append_F(std::string &out, int w, int d, double x)
{
    if (std::isnan(x) || std::isinf(x)) { out += F(w, d, x); return; }
    char buf[400];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed, d);
    if (r.ec != std::errc()) { out += F(w, d, x); return; }
    char *s = buf;
    int n = (int)(r.ptr - buf);
    if (d == 0) s[n++] = '.';   // see F()
    if (n > w) {   // gfortran drops the optional leading zero when the field is too narrow
        if (n >= 2 && s[0] == '0' && s[1] == '.') { s++; n--; }
        else if (n >= 3 && s[0] == '-' && s[1] == '0' && s[2] == '.') { s[1] = '-'; s++; n--; }
    }
    if (n > w) { out.append((size_t)(w > 0 ? w : 1), '*'); return; }
    append_padded(out, w, s, n);
}

// ---- fast path of the map writers (G15.8, E14.6: 0.3 G values per 4096^2 run, where two std::to_chars calls per value
// were the writers' whole cost).  The nd <= 9 leading decimal digits of a come from ONE multiplication by a correctly
// rounded power of ten in 64-bit-mantissa arithmetic (x87 long double): the product is within 2^-62 relative of the exact
// a * 10^n, i.e. within 3e-10 of the exact scaled value (< 1e9), so its rounding to an integer is the correctly rounded
// digit string unless the fraction lies within 1e-7 of one half -- then, and only then (exact ties such as 12345678.5
// included), the caller falls back to std::to_chars.  tests/cpp/fmt_check.cpp compares both paths on tens of millions of
// values (random mantissas and exponents, short decimals, near ties, powers of ten, subnormals).
namespace detail {
struct Pow10Table {
    long double p[701];
    Pow10Table()
    {
        for (int n = -350; n <= 350; n++) {
            char b[16];
            snprintf(b, sizeof b, "1e%d", n);
            p[n + 350] = strtold(b, nullptr);     // correctly rounded
        }
    }
};
inline const long double *pow10l_table()
{
    static const Pow10Table t;
    return t.p + 350;
}
constexpr unsigned long long kPow10u[11] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
                                            100000000ull, 1000000000ull, 10000000000ull};

// a finite and > 0, 1 <= nd <= 9.  On success a rounds to D * 10^(k - nd + 1) with 10^(nd-1) <= D < 10^nd.
inline bool fast_digits(double a, int nd, unsigned long long &D, int &k)
{
    if (sizeof(long double) < 10) return false;          // no 64-bit mantissa here: exact path only
    int e;
    (void)std::frexp(a, &e);                              // a in [2^(e-1), 2^e)
    k = (int)(((long long)(e - 1) * 1292913986LL) >> 32);   // floor((e-1) log10 2): the decimal exponent or one below it
    const long double *p10 = pow10l_table();
    const unsigned long long hi = kPow10u[nd], lo = kPow10u[nd - 1];
    long double scaled = (long double)a * p10[nd - 1 - k];
    if (scaled >= (long double)hi) {
        k++;
        scaled = (long double)a * p10[nd - 1 - k];
    }
    if (!(scaled < (long double)hi)) return false;
    D = (unsigned long long)scaled;                       // truncation
    const long double frac = scaled - (long double)D;
    if (frac > 0.5000001L) D++;
    else if (!(frac < 0.4999999L)) return false;          // too close to call
    if (D >= hi) { D = lo; k++; }                         // 9.9999999|6 -> 10.000000
    return D >= lo;
}

// the digits of D (exactly nd of them) into s
inline void put_digits(char *s, unsigned long long D, int nd)
{
    for (int i = nd - 1; i >= 0; i--) { s[i] = (char)('0' + D % 10); D /= 10; }
}

inline void append_E_digits(std::string &out, int w, int d, bool neg, unsigned long long D, int k)
{
    char s[40];
    int n = 0;
    if (neg) s[n++] = '-';
    s[n++] = '0';
    s[n++] = '.';
    put_digits(s + n, D, d);
    n += d;
    const int ex = k + 1, ax = ex < 0 ? -ex : ex;
    if (ax < 100) {
        s[n++] = 'E';
        s[n++] = ex < 0 ? '-' : '+';
        s[n++] = (char)('0' + ax / 10);
        s[n++] = (char)('0' + ax % 10);
    } else {
        s[n++] = ex < 0 ? '-' : '+';
        s[n++] = (char)('0' + ax / 100);
        s[n++] = (char)('0' + ax / 10 % 10);
        s[n++] = (char)('0' + ax % 10);
    }
    char *q = s;
    if (n > w) {   // drop the leading zero: [-].ddddE+ee
        if (neg) { s[1] = '-'; q = s + 1; }
        else q = s + 1;
        n--;
    }
    if (n > w) { out.append((size_t)(w > 0 ? w : 1), '*'); return; }
    append_padded(out, w, q, n);
}

// F(w).(d - s) of a value whose d significant digits are D and that has s of them (0 <= s <= d) before the point
inline void append_F_digits(std::string &out, int w, int d, int s_int, bool neg, unsigned long long D)
{
    char digits[24], s[40];
    put_digits(digits, D, d);
    int n = 0;
    if (neg) s[n++] = '-';
    if (s_int == 0) s[n++] = '0';
    for (int i = 0; i < s_int; i++) s[n++] = digits[i];
    s[n++] = '.';
    for (int i = s_int; i < d; i++) s[n++] = digits[i];
    char *q = s;
    if (n > w && s_int == 0) {   // the optional leading zero goes first when the field is too narrow
        if (neg) { s[1] = '-'; q = s + 1; }
        else q = s + 1;
        n--;
    }
    if (n > w) { out.append((size_t)(w > 0 ? w : 1), '*'); return; }
    append_padded(out, w, q, n);
}
}   // namespace detail

inline void append_E(std::string &out, int w, int d, double x)
{
    if (std::isnan(x) || std::isinf(x) || d > 40) { out += E(w, d, x); return; }
    if (d >= 1 && d <= 9 && x != 0.0) {
        unsigned long long D;
        int k;
        if (detail::fast_digits(std::fabs(x), d, D, k)) { detail::append_E_digits(out, w, d, std::signbit(x), D, k); return; }
    }
    const bool neg = std::signbit(x);
    const double a = std::fabs(x);
    char buf[64], s[80];
    int ex = 0, n = 0;
    if (neg) s[n++] = '-';
    s[n++] = '0';
    s[n++] = '.';
    if (a == 0.0) {
        for (int k = 0; k < d; k++) s[n++] = '0';
    } else {
        auto r = std::to_chars(buf, buf + sizeof buf, a, std::chars_format::scientific, d - 1);   // D.DDDDe+XX
        if (r.ec != std::errc()) { out += E(w, d, x); return; }
        const char *e = buf;
        while (*e != 'e') e++;
        for (const char *p = buf; p < e; p++)
            if (*p != '.') s[n++] = *p;
        int v = 0;
        for (const char *p = e + 2; p < r.ptr; p++) v = v * 10 + (*p - '0');
        ex = (e[1] == '-' ? -v : v) + 1;
    }
    const int ax = ex < 0 ? -ex : ex;
    if (ax < 100) {
        s[n++] = 'E';
        s[n++] = ex < 0 ? '-' : '+';
        s[n++] = (char)('0' + ax / 10);
        s[n++] = (char)('0' + ax % 10);
    } else {
        s[n++] = ex < 0 ? '-' : '+';
        s[n++] = (char)('0' + ax / 100);
        s[n++] = (char)('0' + ax / 10 % 10);
        s[n++] = (char)('0' + ax % 10);
    }
    char *q = s;
    if (n > w) {   // drop the leading zero: [-].ddddE+ee
        if (neg) { s[1] = '-'; q = s + 1; }
        else q = s + 1;
        n--;
    }
    if (n > w) { out.append((size_t)(w > 0 ? w : 1), '*'); return; }
    append_padded(out, w, q, n);
}

inline void append_G(std::string &out, int w, int d, double x)
{
    if (std::isnan(x) || std::isinf(x) || d > 40) { out += G(w, d, x); return; }
    const double a = std::fabs(x);
    if (a == 0.0) { append_F(out, w - 4, d - 1, x); out.append(4, ' '); return; }
    if (d >= 1 && d <= 9) {
        unsigned long long D;
        int k;
        if (detail::fast_digits(a, d, D, k)) {
            const int s = k + 1;
            if (s < 0 || s > d) detail::append_E_digits(out, w, d, std::signbit(x), D, k);
            else { detail::append_F_digits(out, w - 4, d, s, std::signbit(x), D); out.append(4, ' '); }
            return;
        }
    }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, a, std::chars_format::scientific, d - 1);
    if (r.ec != std::errc()) { out += G(w, d, x); return; }
    const char *e = buf;
    while (*e != 'e') e++;
    int v = 0;
    for (const char *p = e + 2; p < r.ptr; p++) v = v * 10 + (*p - '0');
    const int s = (e[1] == '-' ? -v : v) + 1;
    if (s < 0 || s > d) { append_E(out, w, d, x); return; }
    append_F(out, w - 4, d - s, x);
    out.append(4, ' ');
}

inline std::string rtrim(std::string s)
{
    while (!s.empty() && s.back() == ' ') s.pop_back();
    return s;
}

}   // namespace ffmt
