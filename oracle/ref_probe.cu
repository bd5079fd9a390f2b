/*
 * ref_probe.cu -- runs the REFERENCE's own functions on caller-supplied inputs.  TEST
 * INFRASTRUCTURE ONLY; built into oracle/_ref/ref_probe by oracle/Makefile when
 * /root/reference is present.  No reference source is copied: the translation unit below
 * #includes cudabrot.cu where it lies (path given on the nvcc command line as -DREF_SOURCE=...).
 *
 *   ref_probe tonemap <hist.raw> <w> <h> <gamma> <out.pgm>
 *       host only (no GPU needed): the reference's SetGrayscalePixels (cudabrot.cu:454-468) and
 *       SaveImage (:548-577) on a raw uint32 histogram.
 *   ref_probe orbits <w> <h> <min_re> <max_re> <min_im> <max_im> <max_it> <min_it>
 *                    <samples.f64> <iters_out.i32> <hist_out.raw>
 *       needs a GPU: for every (re, im) pair in samples.f64 runs the reference's InMainCardioid /
 *       InOrder2Bulb / IterateMandelbrot (:284-340) one sample per thread, then IterateAndRecord
 *       (:347-365) for the accepted ones from ONE thread, so the reference's non-atomic `+= 1`
 *       (:312) cannot lose updates.  This is the reference's real SASS on the oracle's sample list.
 */
#define main cudabrot_reference_main
#include REF_SOURCE
#undef main

#include <vector>

__global__ void ProbeClassify(const double *samples, int n, int max_iterations, int *out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double real = samples[2 * k], imag = samples[2 * k + 1];
#ifndef RENDER_BURNING_SHIP  /* as in DrawBuddhabrot, cudabrot.cu:397-399 */
  if (InMainCardioid(real, imag) || InOrder2Bulb(real, imag)) {
    out[k] = -1;
    return;
  }
#endif
  out[k] = IterateMandelbrot(real, imag, max_iterations);
}

__global__ void ProbeRecord(FractalDimensions dimensions, Pixel *data, const double *samples,
                            const int *iters, int n, IterationControl iterations) {
  for (int k = 0; k < n; k++) {
    int it = iters[k];
    if (it < 0) continue;
    if (it >= iterations.max_escape_iterations) continue;
    if (it < iterations.min_escape_iterations) continue;
    IterateAndRecord(samples[2 * k], samples[2 * k + 1], data, &dimensions);
  }
}

static int ReadWhole(const char *path, std::vector<char> &buf) {
  FILE *f = fopen(path, "rb");
  if (!f) return 0;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(sz);
  int ok = sz == 0 || fread(buf.data(), sz, 1, f) == 1;
  fclose(f);
  return ok;
}

static int WriteWhole(const char *path, const void *p, size_t bytes) {
  FILE *f = fopen(path, "wb");
  if (!f) return 0;
  int ok = bytes == 0 || fwrite(p, bytes, 1, f) == 1;
  fclose(f);
  return ok;
}

static int ProbeTonemap(int argc, char **argv) {
  if (argc != 7) return 2;
  memset(&g, 0, sizeof(g));
  SetDefaultCanvas();
  g.dimensions.w = atoi(argv[3]);
  g.dimensions.h = atoi(argv[4]);
  if (!RecomputePixelDeltas()) return 2;
  g.gamma_correction = strtod(argv[5], NULL);
  g.output_image = argv[6];
  std::vector<char> buf;
  if (!ReadWhole(argv[2], buf) || buf.size() != GetImageBufferSize()) {
    printf("bad histogram file\n");
    return 1;
  }
  g.host_buddhabrot = (Pixel *) buf.data();
  std::vector<uint16_t> gray((size_t) g.dimensions.w * g.dimensions.h);
  g.grayscale_image = gray.data();
  SetGrayscalePixels();
  SaveImage();
  return 0;
}

static int ProbeOrbits(int argc, char **argv) {
  if (argc != 13) return 2;
  memset(&g, 0, sizeof(g));
  SetDefaultCanvas();
  g.dimensions.w = atoi(argv[2]);
  g.dimensions.h = atoi(argv[3]);
  g.dimensions.min_real = strtod(argv[4], NULL);
  g.dimensions.max_real = strtod(argv[5], NULL);
  g.dimensions.min_imag = strtod(argv[6], NULL);
  g.dimensions.max_imag = strtod(argv[7], NULL);
  if (!RecomputePixelDeltas()) return 2;
  g.iterations.max_escape_iterations = atoi(argv[8]);
  g.iterations.min_escape_iterations = atoi(argv[9]);
  std::vector<char> buf;
  if (!ReadWhole(argv[10], buf) || (buf.size() % 16) != 0) {
    printf("bad samples file\n");
    return 1;
  }
  int n = (int) (buf.size() / 16);
  double *d_samples = NULL;
  int *d_iters = NULL;
  Pixel *d_hist = NULL;
  CheckCUDAError(cudaSetDevice(0));
  CheckCUDAError(cudaMalloc(&d_samples, buf.size() + 16));
  CheckCUDAError(cudaMalloc(&d_iters, sizeof(int) * (n + 1)));
  CheckCUDAError(cudaMalloc(&d_hist, GetImageBufferSize()));
  CheckCUDAError(cudaMemset(d_hist, 0, GetImageBufferSize()));
  CheckCUDAError(cudaMemcpy(d_samples, buf.data(), buf.size(), cudaMemcpyHostToDevice));
  ProbeClassify<<<(n + 255) / 256, 256>>>(d_samples, n,
    g.iterations.max_escape_iterations, d_iters);
  CheckCUDAError(cudaDeviceSynchronize());
  ProbeRecord<<<1, 1>>>(g.dimensions, d_hist, d_samples, d_iters, n, g.iterations);
  CheckCUDAError(cudaDeviceSynchronize());
  std::vector<int> iters(n);
  std::vector<Pixel> hist((size_t) g.dimensions.w * g.dimensions.h);
  CheckCUDAError(cudaMemcpy(iters.data(), d_iters, sizeof(int) * n, cudaMemcpyDeviceToHost));
  CheckCUDAError(cudaMemcpy(hist.data(), d_hist, GetImageBufferSize(),
    cudaMemcpyDeviceToHost));
  if (!WriteWhole(argv[11], iters.data(), sizeof(int) * n)) return 1;
  if (!WriteWhole(argv[12], hist.data(), GetImageBufferSize())) return 1;
  cudaFree(d_samples);
  cudaFree(d_iters);
  cudaFree(d_hist);
  printf("probe ok: %d samples\n", n);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 2 && strcmp(argv[1], "tonemap") == 0) return ProbeTonemap(argc, argv);
  if (argc >= 2 && strcmp(argv[1], "orbits") == 0) return ProbeOrbits(argc, argv);
  printf("usage: ref_probe tonemap|orbits ... (see oracle/ref_probe.cu)\n");
  return 2;
}
