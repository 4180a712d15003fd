#include "fortran_format.hpp"
#include <random>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <cmath>
#include <vector>
int main(int argc, char **argv){
    std::mt19937_64 rng(1); std::uniform_real_distribution<double> u(-1,1); std::uniform_int_distribution<int> ex(-320,308);
    std::vector<double> vals={0.0,-0.0,1.0,-1.0,0.1,0.099999999,0.0999999999999,9.9999999e7,99999999.5,1e8,1e-5,0.83333333e-16,123456789.0,1e100,-1e-100,1e-310,5e-324,1.7976931348623157e308,0.5,0.05,9.99999995,0.999999995,99999999.4999,1e22,2.5,0.25,0.125};
    for(int i=0;i<300000;i++){ double m=u(rng); int e=ex(rng); vals.push_back(std::ldexp(m, (int)(e*3.32))); if(i%3==0) vals.push_back(m*100); if(i%5==0) vals.push_back(std::round(m*1e6)/1e3);}
    long bad=0;
    for(double x: vals){
        std::string a,b;
        a=ffmt::G(15,8,x); ffmt::append_G(b,15,8,x); if(a!=b){ if(bad<10) printf("G15.8 %.17g: [%s] vs [%s]\n",x,a.c_str(),b.c_str()); bad++; }
        a=ffmt::E(14,6,x); b.clear(); ffmt::append_E(b,14,6,x); if(a!=b){ if(bad<10) printf("E14.6 %.17g: [%s] vs [%s]\n",x,a.c_str(),b.c_str()); bad++; }
        a=ffmt::G(20,12,x); b.clear(); ffmt::append_G(b,20,12,x); if(a!=b){ if(bad<10) printf("G20.12 %.17g: [%s] vs [%s]\n",x,a.c_str(),b.c_str()); bad++; }
        a=ffmt::F(15,6,x); b.clear(); ffmt::append_F(b,15,6,x); if(a!=b){ if(bad<10) printf("F15.6 %.17g: [%s] vs [%s]\n",x,a.c_str(),b.c_str()); bad++; }
    }
    // the map writers' fast path (one 64-bit-mantissa multiplication) against the exact formatters on many more values:
    // random mantissa x exponent, short decimals, near ties of the 8th / 7th digit, powers of ten and their neighbours,
    // subnormals; argv[1] = millions of random values (default 2)
    {
        const long n_rand = (argc > 1 ? atol(argv[1]) : 2) * 1000000L;
        std::uniform_int_distribution<unsigned long long> bits;
        std::uniform_int_distribution<long long> dig8(10000000, 99999999), dig7(1000000, 9999999);
        std::uniform_int_distribution<int> dex(-300, 290);
        long fast_bad = 0, checked = 0;
        auto check = [&](double x) {
            if (!std::isfinite(x)) return;
            std::string a = ffmt::G(15, 8, x), b;
            ffmt::append_G(b, 15, 8, x);
            if (a != b) { if (fast_bad < 10) printf("fast G15.8 %.17g: [%s] vs [%s]\n", x, a.c_str(), b.c_str()); fast_bad++; }
            a = ffmt::E(14, 6, x); b.clear(); ffmt::append_E(b, 14, 6, x);
            if (a != b) { if (fast_bad < 10) printf("fast E14.6 %.17g: [%s] vs [%s]\n", x, a.c_str(), b.c_str()); fast_bad++; }
            checked++;
        };
        for (long i = 0; i < n_rand; i++) {
            double x; unsigned long long u = bits(rng); memcpy(&x, &u, sizeof x); check(x);              // any bit pattern
            if (i % 4 == 0) {
                const int e10 = dex(rng);
                check((double)dig8(rng) * std::pow(10.0, e10));                                           // 8-digit decimals
                check(((double)dig8(rng) + 0.5) * std::pow(10.0, e10));                                   // near ties, 8th digit
                check(((double)dig7(rng) + 0.5) * std::pow(10.0, e10));                                   // near ties, 7th digit
                check(std::nextafter(((double)dig8(rng) + 0.5) * std::pow(10.0, e10), 0.0));
            }
        }
        for (int e10 = -323; e10 <= 308; e10++) {
            const double p = std::pow(10.0, e10);
            for (double x : {p, std::nextafter(p, 0.0), std::nextafter(p, HUGE_VAL), p * 0.99999999, p * 0.999999995, p * 9.9999999,
                             p * 9.99999995, p * 9.999999951, p * 9.9999995, -p, p * 1.00000005, p * 12345678.5, p * 1234567.5})
                check(x);
        }
        for (double x : {12345678.5, 1234567.5, 0.5, 2.5, 99999999.5, 9999999.5, 0.099999995, 0.0999999949, 99999999.49, 1e8, 1e7,
                         12345678.0, 123456789.0, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, -12345678.0, -0.1})
            check(x);
        printf("fast path: %ld values checked, %ld mismatches\n", checked, fast_bad);
        bad += fast_bad;
    }
    // known answers: characters the reference's Fortran build wrote (test/fortran/fixtures/coverage_test_*-orig.*) and the
    // F editing rule that a field always shows its decimal point (G15.8 of 1e7 <= x < 1e8 is F11.0: "12345678.")
    {
        struct { double x; const char *g15_8; const char *e14_6; } known[] = {
            {99.5, "  99.500000    ", "  0.995000E+02"},
            {0.83333333333333e-16, " 0.83333333E-16", "  0.833333E-16"},
            {-0.10271626e-47, "-0.10271626E-47", " -0.102716E-47"},
            {-0.0, " -0.0000000    ", " -0.000000E+00"},
            {0.0, "  0.0000000    ", "  0.000000E+00"},
            {-0.5e-5, "-0.50000000E-05", " -0.500000E-05"},
            {1.0e-4, " 0.10000000E-03", "  0.100000E-03"},
            {12345678.0, "  12345678.    ", "  0.123457E+08"},
            {-12345678.0, " -12345678.    ", " -0.123457E+08"},
            {99999999.5, " 0.10000000E+09", "  0.100000E+09"},
            {0.1, " 0.10000000    ", "  0.100000E+00"},
            {1.0e-100, " 0.10000000E-99", "  0.100000E-99"},
            {1.0e-101, " 0.10000000-100", "  0.100000-100"},
        };
        for (auto &kv : known) {
            std::string b;
            ffmt::append_G(b, 15, 8, kv.x);
            if (b != kv.g15_8 || ffmt::G(15, 8, kv.x) != kv.g15_8) { printf("known G15.8 %.17g: [%s] [%s] want [%s]\n", kv.x, b.c_str(), ffmt::G(15, 8, kv.x).c_str(), kv.g15_8); bad++; }
            b.clear();
            ffmt::append_E(b, 14, 6, kv.x);
            if (b != kv.e14_6 || ffmt::E(14, 6, kv.x) != kv.e14_6) { printf("known E14.6 %.17g: [%s] [%s] want [%s]\n", kv.x, b.c_str(), ffmt::E(14, 6, kv.x).c_str(), kv.e14_6); bad++; }
        }
    }
    printf("values %zu mismatches %ld\n", vals.size(), bad);
    auto t0=std::chrono::steady_clock::now(); size_t tot=0;
    for(double x: vals){ std::string s=ffmt::G(15,8,x); tot+=s.size(); }
    auto t1=std::chrono::steady_clock::now(); std::string buf; 
    for(double x: vals){ ffmt::append_G(buf,15,8,x); if(buf.size()>1<<20) buf.clear(); }
    auto t2=std::chrono::steady_clock::now();
    printf("slow %.1f ns/value fast %.1f ns/value\n", std::chrono::duration<double,std::nano>(t1-t0).count()/vals.size(), std::chrono::duration<double,std::nano>(t2-t1).count()/vals.size());
    return bad!=0;
}
