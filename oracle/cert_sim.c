// Offline experiment (CPU, links oracle/liboracle.so; not part of the product): how many escape-pass iterations would a
// contraction certificate save on never-escaping samples, compared with bit-exact detection?
#include <stdio.h>
// (test infrastructure like the rest of oracle/: never linked or called by the product)
// gcc -O2 -fopenmp -ffp-contract=off -mfma -o _ref/cert_sim cert_sim.c liboracle.so -lm -Wl,-rpath,$PWD
// ./cert_sim 4194304 1e-5 0.9   ->  samples, tolerance of the near-return, multiplier bound
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#include "buddha_oracle.h"
static inline void step(double *re, double *im, double cr, double ci) {
  double t1 = *im * *im; double t2 = fma(*re, *re, -t1); double r2 = *re + *re;
  *im = fma(r2, *im, ci); *re = cr + t2;
}
static int cp_age(unsigned age) { unsigned rest = age & (age - 1u); return rest == 0u || 3u * rest == 2u * age; }
int main(int argc, char **argv) {
  uint64_t n = argc > 1 ? strtoull(argv[1], 0, 0) : (1ull << 22);
  int maxit = 20000; double tol = argc > 2 ? atof(argv[2]) : 1e-5; double lam_max = argc > 3 ? atof(argv[3]) : 0.9;
  const float p3_limit = argc > 4 ? (float)atof(argv[4]) : 0.96f;   // the kernel's period-3 bound on |lambda|^2
  double sum_exact = 0, sum_cert = 0, sum_cert_extra = 0; uint64_t inset = 0, exact_found = 0, cert_found = 0, p3 = 0, bad = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : sum_exact, sum_cert, sum_cert_extra, inset, exact_found, cert_found, p3, bad)
  for (uint64_t s = 0; s < n; s++) {
    double cr, ci; oracle_sample(1337, s, &cr, &ci);
    if (oracle_rejected(cr, ci)) continue;
    // quick: must survive 46 steps
    int it = oracle_escape_iterations(cr, ci, maxit);
    int escapes = it < maxit;
    if (it < 46) continue;
    if (oracle_period3_flag(cr, ci, p3_limit)) { p3++; continue; }
    // simulate deep rounds from step 46
    double re = cr, im = ci; for (int k = 0; k < 46; k++) step(&re, &im, cr, ci);
    double rre = re, rim = im; unsigned age = 0, age_cp = 0; int done = 46; int t_exact = -1, t_cert = -1, cert_cost = 0;
    while (done + 24 <= maxit) {
      if (cp_age(age)) { rre = re; rim = im; age_cp = age; }
      for (int k = 0; k < 24; k++) step(&re, &im, cr, ci);
      done += 24; age++;
      if (!(re * re + im * im <= 4.0)) break;   // escaped (approx; the sample is classified by `escapes`)
      if (t_exact < 0 && memcmp(&re, &rre, 8) == 0 && memcmp(&im, &rim, 8) == 0) t_exact = done;
      if (t_cert < 0 && age > age_cp && fabs(re - rre) < tol && fabs(im - rim) < tol) {
        // near return after P = (age - age_cp) * 24 steps: multiplier along the next P steps
        int P = (int)(age - age_cp) * 24;
        double zr = re, zi = im, lr = 1, li = 0; int ok = 1;
        for (int k = 0; k < P; k++) {
          double nr = 2 * (lr * zr - li * zi), ni = 2 * (lr * zi + li * zr); lr = nr; li = ni;
          step(&zr, &zi, cr, ci);
          if (!(zr * zr + zi * zi <= 4.0)) { ok = 0; break; }
        }
        cert_cost += P * 3;   // ~12 FP64 per certificate step vs 4 per plain step
        if (ok && hypot(lr, li) < lam_max && fabs(zr - re) < tol && fabs(zi - im) < tol) t_cert = done;
      }
      if (t_exact >= 0) break;
    }
    if (escapes) { if (t_cert >= 0) bad++; continue; }
    inset++;
    if (t_exact >= 0) { exact_found++; sum_exact += t_exact; } else sum_exact += maxit;
    if (t_cert >= 0) { cert_found++; sum_cert += t_cert; sum_cert_extra += cert_cost; } else { sum_cert += (t_exact >= 0 ? t_exact : maxit); sum_cert_extra += cert_cost; }
  }
  printf("samples %llu  period-3 flagged %llu  in-set left %llu  exact found %.1f%%  cert found %.1f%%  certified-but-escapes %llu\n",
         (unsigned long long)n, (unsigned long long)p3, (unsigned long long)inset, 100.0 * exact_found / inset, 100.0 * cert_found / inset, (unsigned long long)bad);
  printf("mean iterations per in-set sample: bit-exact %.0f   certificate %.0f (+%.0f step-equivalents of certificate work)\n",
         sum_exact / inset, sum_cert / inset, sum_cert_extra / inset);
  return 0;
}
