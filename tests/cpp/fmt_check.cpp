#include "fortran_format.hpp"
#include <random>
#include <cstdio>
#include <chrono>
int main(){
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
    printf("values %zu mismatches %ld\n", vals.size(), bad);
    auto t0=std::chrono::steady_clock::now(); size_t tot=0;
    for(double x: vals){ std::string s=ffmt::G(15,8,x); tot+=s.size(); }
    auto t1=std::chrono::steady_clock::now(); std::string buf; 
    for(double x: vals){ ffmt::append_G(buf,15,8,x); if(buf.size()>1<<20) buf.clear(); }
    auto t2=std::chrono::steady_clock::now();
    printf("slow %.1f ns/value fast %.1f ns/value\n", std::chrono::duration<double,std::nano>(t1-t0).count()/vals.size(), std::chrono::duration<double,std::nano>(t2-t1).count()/vals.size());
    return bad!=0;
}
