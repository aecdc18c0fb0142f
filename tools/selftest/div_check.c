/* (runs on the CPU, no GPU)  kernels.cuh div_rcp -- a / b from y = RN(1/b) with one (div3) or two (div5, the product's,
 * proven by Markstein's theorem) FMA corrections -- against IEEE division: the reference's constant divisors
 * (C_S^2, 2 C_S^4, 2 C_S^2, typical tau) and random divisors, 3e8 random dividends.
 *   gcc -O2 -ffp-contract=off -mfma -o div_check div_check.c -lm && ./div_check      (expects bad5 = 0 everywhere) */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
static inline double div5(double a, double b, double y){
    double q0=a*y; double r0=fma(-b,q0,a); double q1=fma(r0,y,q0); double r1=fma(-b,q1,a); double q2=fma(r1,y,q1);
    return a==0.0? q0 : q2;
}
static inline double div3(double a, double b, double y){
    double q0=a*y; double r0=fma(-b,q0,a); double q1=fma(r0,y,q0);
    return a==0.0? q0 : q1;
}
static uint64_t s=88172645463325252ULL;
static inline uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s; }
int main(){
    const double C_S=0.57735026919;
    double cs2=C_S*C_S, c4=2.0*C_S*C_S*C_S*C_S, c2=2.0*C_S*C_S;
    double bs[]={cs2,c4,c2,0.6,0.5001,1.9999999, 1.0000000000000002, 0.51, 1.7, 3.0, 0.9999999999999999};
    long bad3=0,bad5=0; long N=200000000;
    for(int k=0;k<11;k++){
        double b=bs[k], y=1.0/b; long b3=0,b5=0;
        for(long i=0;i<N/11;i++){
            uint64_t m=rnd(); int e=(int)(rnd()%40)-30; 
            uint64_t bits=((uint64_t)(1023+e)<<52)|(m>>12); if(rnd()&1) bits|=1ULL<<63;
            double a; memcpy(&a,&bits,8);
            double t=a/b;
            if(div3(a,b,y)!=t) b3++;
            if(div5(a,b,y)!=t) b5++;
        }
        printf("b=%.17g bad3=%ld bad5=%ld\n",b,b3,b5); bad3+=b3;bad5+=b5;
    }
    /* variable b */
    long b3=0,b5=0;
    for(long i=0;i<100000000;i++){
        uint64_t m=rnd(); uint64_t bits=((uint64_t)(1023+(int)(rnd()%6)-3)<<52)|(m>>12); double a; memcpy(&a,&bits,8);
        m=rnd(); bits=((uint64_t)(1023+(int)(rnd()%4)-2)<<52)|(m>>12); double b; memcpy(&b,&bits,8);
        double y=1.0/b, t=a/b;
        if(div3(a,b,y)!=t) b3++;
        if(div5(a,b,y)!=t) b5++;
    }
    printf("variable b: bad3=%ld bad5=%ld\n",b3,b5);
    return 0;
}
