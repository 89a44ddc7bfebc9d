/* A plain-C client of the C ABI (include/ppb200.h): reads a portrait, its model and the channel
 * frequencies from a raw file written by the test, runs pp_set_model + pp_fit_batch and prints
 * the fitted parameters.  Built with gcc and run by tests/test_gpu_c_client.py; no Python, no
 * torch on this side of the boundary.
 *
 *   abi_fit <file> <nchan> <nbin> <nsub> <P>
 * file layout: freqs f64[nchan], model f32[nchan*nbin], data f32[nsub*nchan*nbin]            */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ppb200.h"

#define DIE(msg) do { fprintf(stderr, "%s: %s\n", msg, pp_last_error()); return 1; } while (0)

int main(int argc, char** argv) {
  if (argc != 6) { fprintf(stderr, "usage: abi_fit file nchan nbin nsub P\n"); return 2; }
  const int nchan = atoi(argv[2]), nbin = atoi(argv[3]), nsub = atoi(argv[4]);
  const double period = atof(argv[5]);
  FILE* fh = fopen(argv[1], "rb");
  if (!fh) { perror("open"); return 2; }
  double* freqs = malloc(sizeof(double) * nchan);
  float* model = malloc(sizeof(float) * (size_t)nchan * nbin);
  float* data = malloc(sizeof(float) * (size_t)nsub * nchan * nbin);
  if (fread(freqs, sizeof(double), nchan, fh) != (size_t)nchan ||
      fread(model, sizeof(float), (size_t)nchan * nbin, fh) != (size_t)nchan * nbin ||
      fread(data, sizeof(float), (size_t)nsub * nchan * nbin, fh) != (size_t)nsub * nchan * nbin) {
    fprintf(stderr, "short read\n"); return 2;
  }
  fclose(fh);
  if (pp_abi_version() != PPB200_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 3; }

  pp_plan_t* plan = NULL;
  if (pp_plan_create(nchan, nbin, 0, &plan)) DIE("pp_plan_create");
  if (pp_set_model(plan, model, freqs)) DIE("pp_set_model");

  double* P = malloc(sizeof(double) * nsub);
  for (int i = 0; i < nsub; ++i) P[i] = period;
  pp_fit_args_t a;
  memset(&a, 0, sizeof a);                 /* NULL = defaults: FFTFIT guess, measured noise, all channels */
  a.data = data; a.nsub = nsub; a.semantics = PP_SEM_FIT_PORTRAIT_FULL; a.P = P;
  a.fit_flags[0] = 1; a.fit_flags[1] = 1;
  pp_fit_out_t o;
  memset(&o, 0, sizeof o);                 /* NULL members are not written */
  double* params = malloc(sizeof(double) * nsub * 5);
  double* perrs = malloc(sizeof(double) * nsub * 5);
  double* chi2 = malloc(sizeof(double) * nsub);
  double* nu_out = malloc(sizeof(double) * nsub * 3);
  int32_t* rc = malloc(sizeof(int32_t) * nsub);
  int32_t* lag = malloc(sizeof(int32_t) * nsub);
  o.params = params; o.param_errs = perrs; o.chi2 = chi2; o.nu_out = nu_out; o.return_code = rc; o.lag_index = lag;
  if (pp_fit_batch(plan, &a, &o)) DIE("pp_fit_batch");
  for (int i = 0; i < nsub; ++i)
    printf("%d %.17g %.17g %.17g %.17g %.17g %.17g %d %d\n", i, params[5 * i], perrs[5 * i], params[5 * i + 1],
           perrs[5 * i + 1], chi2[i], nu_out[3 * i], (int)rc[i], (int)lag[i]);

  /* argument errors come back as a negative status with a message, never as a crash */
  a.nsub = 0;
  if (pp_fit_batch(plan, &a, &o) >= 0) { fprintf(stderr, "nsub = 0 was accepted\n"); return 4; }
  if (strlen(pp_last_error()) == 0) { fprintf(stderr, "no error text\n"); return 4; }
  pp_plan_destroy(plan);
  free(freqs); free(model); free(data); free(P); free(params); free(perrs); free(chi2); free(nu_out); free(rc); free(lag);
  return 0;
}
