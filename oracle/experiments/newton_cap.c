/* newton_cap.c -- TEST INFRASTRUCTURE / evidence for DESIGN.md section 3.  Not the product.
 *
 * How often does the REFERENCE's Kepler solver (evidence/rvmodel/trueanomaly.c:15-34: Newton from
 * E = M, stop at |dE| <= 1e-4, abort the call at 10000 iterations) fail to converge, in its own
 * arithmetic (C library sin/cos, IEEE division)?  M is drawn like the model forms it: 2 pi / P *
 * (t - epoch) + M0 with P ~ Jeffreys(1, 1000) d, |t - epoch| < 2500 d.
 *
 *     gcc -O2 -ffp-contract=off -o /tmp/newton_cap oracle/experiments/newton_cap.c -lm
 *     /tmp/newton_cap > profiles/r2_newton_cap.txt
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
int main(){
  const double tol=1e-4; 
  srand48(777);
  for (int band=0; band<6; band++){
    double elo = 0.95+0.01*band, ehi=elo+0.01; if (band==4){elo=0.99;ehi=0.99;} if(band==5){elo=0.985;ehi=0.990;}
    long long n=0, over1k=0, cap=0; int maxit=0; double wM=0,we=0;
    for (long long i=0;i<20000000LL;i++){
      double e = elo + (ehi-elo)*drand48();
      double P = exp(drand48()*log(1000.0)); double t = (drand48()-0.5)*5000.0; double M = 6.283185307179586/P*t + drand48()*6.283185307179586;
      double E=M,E0; int it=0;
      do { E0=E; double ff=E-e*sin(E)-M, dff=1-e*cos(E); E=E0-ff/dff; it++; if(it>=10000) break;} while (fabs(E-E0)>tol);
      n++; if(it>1000) over1k++; if(it>=10000){cap++; wM=M; we=e;} if(it>maxit){maxit=it;}
    }
    printf("e in [%.3f,%.3f]: n=%lld maxit=%d >1000: %lld cap(>=10000): %lld  example e=%.17g M=%.17g\n", elo, ehi, n, maxit, over1k, cap, we, wM);
    fflush(stdout);
  }
}
