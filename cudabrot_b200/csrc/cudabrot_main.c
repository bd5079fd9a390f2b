/*
 * cudabrot_main.c -- the drop-in command line, plain C on top of the C ABI (include/buddha.h).
 *
 * Keeps the reference's flag surface, defaults, validation order, messages, the headerless uint32
 * -s file and the 16-bit big-endian PGM (cudabrot.cu:579-754 flags, :762-791 main flow, :215-280
 * -s, :548-577 PGM).  The hot path itself lives in libbuddha.so.
 *
 * Extensions (absent => reference behaviour):
 *   --samples N       render exactly N candidates (reproducible) instead of running for -t seconds
 *   --first-sample K  first Philox sample index (default 0, or the -s sidecar's cursor)
 *   --seed S          Philox key (default 1337 = the reference's DEFAULT_RNG_SEED, cudabrot.cu:37)
 *   --gpus G          use devices d..d+G-1, disjoint sample ranges, one ncclReduce at the end
 *   --no-shortcut     disable the exact periodicity shortcut (identical output, slower)
 *   --burning-ship    the reference's compile-time RENDER_BURNING_SHIP variant (cudabrot.cu:15-17)
 *   --channels M1:C1,M2:C2[,...]  fused multi-channel render: one pass over the candidates, one
 *                     histogram / -s file / PGM per (-m, -c) pair (what the three runs of
 *                     generate_hires_color_image.sh:27-59 produce); files get a .chK suffix
 *   --color rgb|hsl   with three or more channels: also write a 16-bit colour PPM combined on the
 *                     GPU from channels 0, 1, 2 (red/green/blue, or hue/saturation/lightness: the
 *                     step generate_hires_color_image.sh:61-71 leaves to external tools);
 *                     --hue-adjust X shifts the hue (the script uses 0.3)
 * With -s FILE the next sample index is kept in FILE.cursor so a resumed run continues the stream
 * instead of replaying it (the reference re-seeds with 1337 and replays).
 */
#include <errno.h>
#include <pthread.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/buddha.h"

#define MAX_GPUS 16
#define REFERENCE_PASS_SAMPLES 13107200ull /* 512 blocks * 512 threads * 50, cudabrot.cu:20,23,34 */

static struct {
  buddha_params params;
  const char *output_image;
  const char *inprogress_file;
  double seconds_to_run;
  double gamma_correction;
  int gpus;
  int have_samples;
  uint64_t samples;
  int have_first;
  uint64_t first_sample;
  int color_mode;          /* -1 = none, BUDDHA_COMBINE_RGB, BUDDHA_COMBINE_HSL */
  double hue_adjust;
  volatile int quit_signal_received;
  buddha_ctx *ctx[MAX_GPUS];
  uint32_t *host_buddhabrot;
  uint16_t *grayscale_image;
} g;

static void Cleanup(void) {
  for (int i = 0; i < MAX_GPUS; i++) {
    buddha_destroy(g.ctx[i]);
    g.ctx[i] = NULL;
  }
  free(g.host_buddhabrot);
  free(g.grayscale_image);
  g.host_buddhabrot = NULL;
  g.grayscale_image = NULL;
}

static double CurrentSeconds(void) {
  struct timespec ts;
  if (clock_gettime(CLOCK_REALTIME, &ts) != 0) {
    printf("Error getting time.\n");
    exit(1);
  }
  return ((double) ts.tv_sec) + (((double) ts.tv_nsec) / 1e9);
}

/* The reference prints "CUDA error ..." and exits with 1 (cudabrot.cu:134-141). */
static void Check(int rc, buddha_ctx *ctx, const char *what) {
  if (rc == BUDDHA_OK) return;
  printf("%s failed: %s\n", what, buddha_last_error(ctx));
  Cleanup();
  exit(1);
}

static uint64_t ImageBufferSize(void) {
  return ((uint64_t) g.params.width) * ((uint64_t) g.params.height) * sizeof(uint32_t);
}

static int Channels(void) { return g.params.n_channels > 1 ? (int) g.params.n_channels : 1; }

/* File of channel k: `base` itself for a plain render, else base with ".ch<k>" appended (-s) or
 * inserted before the extension (-o: out.pgm -> out.ch0.pgm). */
static void ChannelFileName(char *dst, size_t n, const char *base, int k, int before_extension) {
  if (Channels() == 1) {
    snprintf(dst, n, "%s", base);
    return;
  }
  const char *dot = before_extension ? strrchr(base, '.') : NULL;
  if (dot && !strchr(dot, '/')) snprintf(dst, n, "%.*s.ch%d%s", (int) (dot - base), base, k, dot);
  else snprintf(dst, n, "%s.ch%d", base, k);
}

static void PrintUsage(char *program_name) {
  printf("Usage: %s [options]\n\n", program_name);
  printf("Options may be one or more of the following:\n"
    "  --help: Prints these instructions.\n"
    "  -d <device number>: Sets which GPU to use. Defaults to GPU 0.\n"
    "  -o <output file name>: If provided, the rendered image will be saved\n"
    "     to a .pgm file with the given name. Otherwise, saves the image\n"
    "     to output.pgm.\n"
    "  -m <max escape iterations>: The maximum number of iterations to use\n"
    "     before giving up on seeing whether a point escapes.\n"
    "  -c <min escape iterations>: If a point escapes before this number of\n"
    "     iterations, it will be ignored.\n"
    "  -g <gamma correction>: A gamma-correction value to use on the\n"
    "     resulting image. If negative, no gamma correction will occur.\n"
    "  -t <seconds to run>: A number of seconds to run the calculation for.\n"
    "     Defaults to 10.0. If negative, the program will run continuously\n"
    "     and will terminate (saving the image) when it receives a SIGINT.\n"
    "  -w <width>: The width of the output image, in pixels. Defaults to\n"
    "     1000.\n"
    "  -h <height>: The height of the output image, in pixels. Defaults to\n"
    "     1000.\n"
    "  -s <save/load file>: If provided, this gives a file name into which\n"
    "     the rendering buffer will be saved, for future continuation.\n"
    "     If the program is loaded and the file exists, the buffer will be\n"
    "     filled with the contents of the file, but the dimensions must\n"
    "     match. Note that this file may be huge for high-resolution images.\n"
    "\n"
    "The following settings control the location of the output image on the\n"
    "complex plane, but samples are always drawn from the entire Mandelbrot-\n"
    "set domain (-2-2i to 2+2i). So these settings can be used to save\n"
    "memory or \"crop\" the output, but won't otherwise speed up rendering:\n"
    "  --min-real <min real>: The minimum value along the real axis to\n"
    "             include in the output image. Defaults to -2.0.\n"
    "  --max-real <max real>: The maximum value along the real axis to\n"
    "             include in the output image. Defaults to 2.0.\n"
    "  --min-imag <min imag>: The minimum value along the imaginary axis to\n"
    "             include in the output image. Defaults to -2.0.\n"
    "  --max-imag <max imag>: The maximum value along the imaginary axis to\n"
    "             include in the output image. Defaults to 2.0.\n"
    "\n"
    "Extensions of this build (the reference has none of these):\n"
    "  --samples <N>: Render exactly N candidate samples instead of running\n"
    "     for -t seconds; the result is reproducible bit for bit.\n"
    "  --first-sample <K>: Index of the first sample in the Philox stream.\n"
    "  --seed <S>: Philox key. Defaults to 1337.\n"
    "  --gpus <G>: Use G GPUs starting at -d; histograms are summed at the end.\n"
    "  --no-shortcut: Disable the exact periodicity shortcut (same output).\n"
    "  --burning-ship: Render the burning ship fractal (RENDER_BURNING_SHIP in the reference).\n"
    "  --channels <m1:c1,m2:c2,...>: Render up to 4 (-m, -c) pairs in one pass; the -s and -o\n"
    "      files get a .ch<k> suffix.\n"
    "  --color <rgb|hsl>: With 3 or more channels, also write a colour .ppm combined from\n"
    "      channels 0, 1, 2. --hue-adjust <x> shifts the hue of the hsl mode.\n"
    "");
  exit(0);
}

static int ParseIntArg(int argc, char **argv, int index) {
  char *end = NULL;
  if ((index + 1) >= argc) {
    printf("Argument %s needs a value.\n", argv[index]);
    PrintUsage(argv[0]);
  }
  long v = strtol(argv[index + 1], &end, 10);
  if ((*end != 0) || (argv[index + 1][0] == 0)) {
    printf("Invalid number given to argument %s: %s\n", argv[index], argv[index + 1]);
    PrintUsage(argv[0]);
  }
  return (int) v;
}

static uint64_t ParseU64Arg(int argc, char **argv, int index) {
  char *end = NULL;
  if ((index + 1) >= argc) {
    printf("Argument %s needs a value.\n", argv[index]);
    PrintUsage(argv[0]);
  }
  errno = 0;
  unsigned long long v = strtoull(argv[index + 1], &end, 0);
  if ((*end != 0) || (argv[index + 1][0] == 0) || (argv[index + 1][0] == '-') || errno) {
    printf("Invalid number given to argument %s: %s\n", argv[index], argv[index + 1]);
    PrintUsage(argv[0]);
  }
  return (uint64_t) v;
}

static double ParseDoubleArg(int argc, char **argv, int index) {
  char *end = NULL;
  if ((index + 1) >= argc) {
    printf("Argument %s needs a value.\n", argv[index]);
    PrintUsage(argv[0]);
  }
  double v = strtod(argv[index + 1], &end);
  if ((*end != 0) || (argv[index + 1][0] == 0)) {
    printf("Invalid number given to argument %s: %s\n", argv[index], argv[index + 1]);
    PrintUsage(argv[0]);
  }
  return v;
}

/* Every canvas flag re-validates at once, like the reference (cudabrot.cu:704-748): the order of
 * flags matters, and an invalid canvas prints the message, the usage text and exits with 0. */
static void RevalidateCanvas(char *program_name) {
  const char *why = NULL;
  if (buddha_validate_canvas(&g.params, NULL, NULL, &why) != BUDDHA_OK) {
    printf("%s\n", why);
    PrintUsage(program_name);
  }
}

static void ParseArguments(int argc, char **argv) {
  for (int i = 1; i < argc; i++) {
    const char *a = argv[i];
    if (strcmp(a, "--help") == 0) PrintUsage(argv[0]);
    if (strcmp(a, "-d") == 0) { g.params.device = ParseIntArg(argc, argv, i++); continue; }
    if (strcmp(a, "-o") == 0) {
      if ((i + 1) >= argc) { printf("Missing output file name.\n"); PrintUsage(argv[0]); }
      g.output_image = argv[++i];
      continue;
    }
    if (strcmp(a, "-s") == 0) {
      if ((i + 1) >= argc) { printf("Missing in-progress buffer file name.\n"); PrintUsage(argv[0]); }
      g.inprogress_file = argv[++i];
      continue;
    }
    if (strcmp(a, "-m") == 0) {
      g.params.max_iterations = ParseIntArg(argc, argv, i++);
      if (g.params.max_iterations > 60000) {
        printf("Warning: Using a high number of iterations may cause the "
          "program respond slowly to Ctrl+C or time running out.\n");
      }
      continue;
    }
    if (strcmp(a, "-c") == 0) { g.params.min_iterations = ParseIntArg(argc, argv, i++); continue; }
    if (strcmp(a, "-w") == 0) {
      g.params.width = ParseIntArg(argc, argv, i++);
      RevalidateCanvas(argv[0]);
      continue;
    }
    if (strcmp(a, "-h") == 0) {
      g.params.height = ParseIntArg(argc, argv, i++);
      RevalidateCanvas(argv[0]);
      continue;
    }
    if (strcmp(a, "-g") == 0) { g.gamma_correction = ParseDoubleArg(argc, argv, i++); continue; }
    if (strcmp(a, "-t") == 0) { g.seconds_to_run = ParseDoubleArg(argc, argv, i++); continue; }
    if (strcmp(a, "--min-real") == 0) {
      g.params.min_real = ParseDoubleArg(argc, argv, i++);
      RevalidateCanvas(argv[0]);
      continue;
    }
    if (strcmp(a, "--max-real") == 0) {
      g.params.max_real = ParseDoubleArg(argc, argv, i++);
      RevalidateCanvas(argv[0]);
      continue;
    }
    if (strcmp(a, "--min-imag") == 0) {
      g.params.min_imag = ParseDoubleArg(argc, argv, i++);
      RevalidateCanvas(argv[0]);
      continue;
    }
    if (strcmp(a, "--max-imag") == 0) {
      g.params.max_imag = ParseDoubleArg(argc, argv, i++);
      RevalidateCanvas(argv[0]);
      continue;
    }
    if (strcmp(a, "--samples") == 0) {
      g.samples = ParseU64Arg(argc, argv, i++);
      g.have_samples = 1;
      continue;
    }
    if (strcmp(a, "--first-sample") == 0) {
      g.first_sample = ParseU64Arg(argc, argv, i++);
      g.have_first = 1;
      continue;
    }
    if (strcmp(a, "--seed") == 0) { g.params.seed = ParseU64Arg(argc, argv, i++); continue; }
    if (strcmp(a, "--gpus") == 0) {
      g.gpus = ParseIntArg(argc, argv, i++);
      if (g.gpus < 1 || g.gpus > MAX_GPUS) {
        printf("--gpus must be between 1 and %d.\n", MAX_GPUS);
        PrintUsage(argv[0]);
      }
      continue;
    }
    if (strcmp(a, "--no-shortcut") == 0) { g.params.flags |= BUDDHA_F_NO_SHORTCUT; continue; }
    if (strcmp(a, "--burning-ship") == 0) { g.params.flags |= BUDDHA_F_BURNING_SHIP; continue; }
    if (strcmp(a, "--channels") == 0) {
      if ((i + 1) >= argc) {
        printf("Argument %s needs a value.\n", a);
        PrintUsage(argv[0]);
      }
      const char *v = argv[++i];
      uint32_t n = 0;
      while (*v && n < BUDDHA_MAX_CHANNELS) {
        char *end = NULL;
        long m = strtol(v, &end, 10);
        if (end == v || *end != ':') break;
        v = end + 1;
        long c = strtol(v, &end, 10);
        if (end == v || (*end != ',' && *end != 0)) break;
        g.params.channel_max[n] = (int32_t) m;
        g.params.channel_min[n] = (int32_t) c;
        n++;
        v = (*end == ',') ? end + 1 : end;
      }
      if (*v || n < 2) {
        printf("--channels needs 2 to %d max:min pairs, e.g. 100:20,1000:20,20000:20\n",
          BUDDHA_MAX_CHANNELS);
        PrintUsage(argv[0]);
      }
      g.params.n_channels = n;
      continue;
    }
    if (strcmp(a, "--color") == 0) {
      if ((i + 1) >= argc) { printf("Argument %s needs a value.\n", a); PrintUsage(argv[0]); }
      const char *v = argv[++i];
      if (strcmp(v, "rgb") == 0) g.color_mode = BUDDHA_COMBINE_RGB;
      else if (strcmp(v, "hsl") == 0) g.color_mode = BUDDHA_COMBINE_HSL;
      else { printf("--color needs rgb or hsl.\n"); PrintUsage(argv[0]); }
      continue;
    }
    if (strcmp(a, "--hue-adjust") == 0) { g.hue_adjust = ParseDoubleArg(argc, argv, i++); continue; }
    printf("Invalid argument: %s\n", a);
    PrintUsage(argv[0]);
  }
  if (g.color_mode >= 0 && g.params.n_channels < 3) {
    printf("--color needs --channels with at least three max:min pairs.\n");
    PrintUsage(argv[0]);
  }
}

static void SignalHandler(int signal_number) {
  g.quit_signal_received = 1;
  printf("Signal %d received, waiting for current pass to finish...\n", signal_number);
}

/* ---- -s file (cudabrot.cu:215-280) plus the cursor sidecar ------------------------------ */

static void SidecarName(char *dst, size_t n) { snprintf(dst, n, "%s.cursor", g.inprogress_file); }

static void LoadInProgressBuffer(void) {
  if (!g.inprogress_file) return;
  uint64_t expected_size = ImageBufferSize();
  if (Channels() > 1) {
    /* all channel files or none: loading a partial set would restart the sample stream and the
     * save at the end would overwrite the channels that do exist */
    int present = 0;
    char missing[4096] = "";
    for (int k = 0; k < Channels(); k++) {
      char file[4096];
      ChannelFileName(file, sizeof(file), g.inprogress_file, k, 0);
      if (access(file, F_OK) == 0) present++;
      else if (!missing[0]) snprintf(missing, sizeof(missing), "%s", file);
    }
    if (present != 0 && present != Channels()) {
      printf("Only %d of the %d channel files of %s exist (%s is missing). Not overwriting them.\n",
        present, Channels(), g.inprogress_file, missing);
      Cleanup();
      exit(1);
    }
  }
  for (int k = 0; k < Channels(); k++) {
    char file[4096];
    ChannelFileName(file, sizeof(file), g.inprogress_file, k, 0);
    FILE *f = fopen(file, "rb");
    printf("Loading previous image state from %s.\n", file);
    if (!f) {
      if (errno == ENOENT) {
        printf("File %s doesn't exist yet. Not loading.\n", file);
        return;
      }
      printf("Failed opening %s: %s\n", file, strerror(errno));
      Cleanup();
      exit(1);
    }
    int64_t file_size = -1;
    if (fseek(f, 0, SEEK_END) == 0) file_size = ftell(f);
    if (file_size < 0 || fseek(f, 0, SEEK_SET) != 0) {
      printf("Failed reading file size: %s\n", strerror(errno));
      fclose(f);
      Cleanup();
      exit(1);
    }
    if ((uint64_t) file_size != expected_size) {
      printf("The size of %s doesn't match the expected size of %lu bytes.\n",
        file, (unsigned long) expected_size);
      fclose(f);
      Cleanup();
      exit(1);
    }
    if (fread((char *) g.host_buddhabrot + (uint64_t) k * expected_size, expected_size, 1, f) != 1) {
      printf("Failed reading %s: %s\n", file, strerror(errno));
      fclose(f);
      Cleanup();
      exit(1);
    }
    fclose(f);
  }
  /* the loaded counts go to GPU 0 only; the other GPUs start from zero and are summed in */
  Check(buddha_load_histogram(g.ctx[0], g.host_buddhabrot,
    (size_t) g.params.width * g.params.height * Channels()), g.ctx[0], "Loading the histogram");

  /* continue the Philox stream where the previous run stopped, if it left a cursor */
  if (!g.have_first) {
    char name[4096];
    SidecarName(name, sizeof(name));
    FILE *c = fopen(name, "r");
    if (c) {
      unsigned long long seed = 0, next = 0;
      if (fscanf(c, "seed %llu next %llu", &seed, &next) == 2 && seed == g.params.seed) {
        g.first_sample = next;
        printf("Continuing the sample stream at index %llu.\n", next);
      }
      fclose(c);
    }
  }
}

static void SaveInProgressBuffer(uint64_t next_sample) {
  if (!g.inprogress_file) return;
  for (int k = 0; k < Channels(); k++) {
    char file[4096];
    ChannelFileName(file, sizeof(file), g.inprogress_file, k, 0);
    printf("Saving in-progress buffer to %s.\n", file);
    FILE *f = fopen(file, "wb");
    if (!f) {
      printf("Failed opening %s: %s\n", file, strerror(errno));
      Cleanup();
      exit(1);
    }
    if (fwrite((char *) g.host_buddhabrot + (uint64_t) k * ImageBufferSize(), ImageBufferSize(), 1,
               f) != 1) {
      printf("Failed writing data to %s: %s\n", file, strerror(errno));
      fclose(f);
      Cleanup();
      exit(1);
    }
    fclose(f);
  }
  char name[4096];
  SidecarName(name, sizeof(name));
  FILE *c = fopen(name, "w");
  if (c) {
    fprintf(c, "seed %llu next %llu\n", (unsigned long long) g.params.seed,
      (unsigned long long) next_sample);
    fclose(c);
  }
}

/* SaveImage (cudabrot.cu:548-577).  The pixels arrive big-endian from the tone-map kernel. */
static void SaveImage(const char *name) {
  size_t pixel_count = (size_t) g.params.width * g.params.height;
  FILE *output = fopen(name, "wb");
  if (!output) {
    printf("Failed opening output image.\n");
    return;
  }
  if (fprintf(output, "P5\n%d %d\n%d\n", g.params.width, g.params.height, 0xffff) <= 0) {
    printf("Failed writing pgm header.\n");
    fclose(output);
    return;
  }
  if (!fwrite(g.grayscale_image, pixel_count * sizeof(uint16_t), 1, output)) {
    printf("Failed writing pixel data.\n");
    fclose(output);
    return;
  }
  fclose(output);
}

/* SetGrayscalePixels (cudabrot.cu:454-468) for one channel, into g.grayscale_image. */
static void ToneMapChannel(int k) {
  uint32_t max = 0;
  double scale = 0;
  Check(buddha_tonemap_channel_u16(g.ctx[0], k, g.gamma_correction, 1, g.grayscale_image,
    (size_t) g.params.width * g.params.height, &max, &scale), g.ctx[0], "Tone-mapping");
  printf("Max value: %lu, scale: %f\n", (unsigned long) max, scale);
}

/* Colour PPM from channels 0, 1, 2, combined on the GPU (generate_hires_color_image.sh:61-71). */
static void SaveColorImage(void) {
  size_t pixel_count = (size_t) g.params.width * g.params.height;
  const int channels[3] = {0, 1, 2};
  uint32_t max[3];
  char name[4096];
  const char *dot = strrchr(g.output_image, '.');
  if (dot && !strchr(dot, '/')) snprintf(name, sizeof(name), "%.*s.color.ppm", (int) (dot - g.output_image), g.output_image);
  else snprintf(name, sizeof(name), "%s.color.ppm", g.output_image);
  uint16_t *rgb = (uint16_t *) malloc(pixel_count * 3 * sizeof(uint16_t));
  if (!rgb) {
    printf("Failed allocating the colour image.\n");
    return;
  }
  Check(buddha_combine_rgb_u16(g.ctx[0], channels, g.gamma_correction, g.color_mode, g.hue_adjust, 1,
    rgb, pixel_count, max), g.ctx[0], "Combining the colour image");
  FILE *output = fopen(name, "wb");
  if (!output) {
    printf("Failed opening output image.\n");
    free(rgb);
    return;
  }
  if (fprintf(output, "P6\n%d %d\n%d\n", g.params.width, g.params.height, 0xffff) <= 0 ||
      !fwrite(rgb, pixel_count * 3 * sizeof(uint16_t), 1, output)) {
    printf("Failed writing the colour image.\n");
  } else {
    printf("Colour image (%s of channels 0, 1, 2) saved: %s\n",
      g.color_mode == BUDDHA_COMBINE_HSL ? "hue/saturation/lightness" : "red/green/blue", name);
  }
  fclose(output);
  free(rgb);
}

/* ---- rendering --------------------------------------------------------------------------- */

typedef struct {
  int index;
  uint64_t first, count;  /* count == 0: time-bounded */
  uint64_t done, passes;
  int rc;
} Worker;

static void *WorkerMain(void *arg) {
  Worker *w = (Worker *) arg;
  buddha_ctx *ctx = g.ctx[w->index];
  if (w->count) {
    w->rc = buddha_render_samples(ctx, w->first, w->count);
    w->done = w->count;
    w->passes = 1;
  } else {
    w->rc = buddha_render_seconds(ctx, g.seconds_to_run, &g.quit_signal_received, w->first,
      &w->done, &w->passes);
  }
  return NULL;
}

/* Returns the next unused sample index. */
static uint64_t RenderImage(void) {
  Worker workers[MAX_GPUS];
  pthread_t threads[MAX_GPUS];
  double start_seconds;
  uint64_t total = 0;
  printf("Calculating Buddhabrot.\n");
  if (g.have_samples) {
    printf("Rendering %llu samples.\n", (unsigned long long) g.samples);
  } else if (g.seconds_to_run < 0) {
    printf("Press ctrl+C to finish.\n");
  } else {
    printf("Running for %.03f seconds.\n", g.seconds_to_run);
  }
  start_seconds = CurrentSeconds();
  for (int i = 0; i < g.gpus; i++) {
    Worker *w = workers + i;
    memset(w, 0, sizeof(*w));
    w->index = i;
    if (g.have_samples) {
      /* contiguous split: the sum over GPUs equals the 1-GPU render of the same range */
      uint64_t per = g.samples / (uint64_t) g.gpus, extra = g.samples % (uint64_t) g.gpus;
      uint64_t lo = per * (uint64_t) i + ((uint64_t) i < extra ? (uint64_t) i : extra);
      uint64_t n = per + ((uint64_t) i < extra ? 1 : 0);
      w->first = g.first_sample + lo;
      w->count = n;
      if (n == 0) { w->done = 0; continue; }
    } else {
      /* time-bounded: GPU i owns the disjoint index range starting at i * 2^56 */
      w->first = g.first_sample + ((uint64_t) i << 56);
    }
  }
  for (int i = 1; i < g.gpus; i++) pthread_create(threads + i, NULL, WorkerMain, workers + i);
  WorkerMain(workers);
  for (int i = 1; i < g.gpus; i++) pthread_join(threads[i], NULL);
  for (int i = 0; i < g.gpus; i++) {
    Check(workers[i].rc, g.ctx[i], "Rendering");
    total += workers[i].done;
  }
  if (g.gpus > 1) Check(buddha_merge(g.ctx, g.gpus, 0), g.ctx[0], "Merging the histograms");
  Check(buddha_read_histogram(g.ctx[0], g.host_buddhabrot,
    (size_t) g.params.width * g.params.height * Channels()), g.ctx[0], "Reading the histogram");
  double seconds = CurrentSeconds() - start_seconds;
  /* one reference pass = 512*512*50 candidates; keeps "passes * 13107200 / seconds" meaningful */
  printf("%d Buddhabrot passes took %f seconds.\n",
    (int) ((total + REFERENCE_PASS_SAMPLES - 1) / REFERENCE_PASS_SAMPLES), seconds);
  printf("%llu candidate samples on %d GPU(s), %.4e samples/s.\n", (unsigned long long) total,
    g.gpus, (double) total / seconds);
  if (Channels() == 1) ToneMapChannel(0);
  /* next unused index: all GPUs' ranges restart from it (plus i * 2^56) on a resumed run */
  uint64_t advance = g.have_samples ? g.samples : 0;
  if (!g.have_samples) {
    for (int i = 0; i < g.gpus; i++) if (workers[i].done > advance) advance = workers[i].done;
  }
  return g.first_sample + advance;
}

int main(int argc, char **argv) {
  memset(&g, 0, sizeof(g));
  buddha_default_params(&g.params);
  g.output_image = "output.pgm";
  g.seconds_to_run = 10.0;
  g.gamma_correction = 1.0;
  g.gpus = 1;
  g.color_mode = -1;
  ParseArguments(argc, argv);
  if (signal(SIGINT, SignalHandler) == SIG_ERR) {
    printf("Failed setting signal handler.\n");
    return 1;
  }
  if (Channels() > 1) {  /* the line below names the widest channel, as the fused pass runs */
    for (int k = 0; k < Channels(); k++)
      if (k == 0 || g.params.channel_max[k] > g.params.max_iterations)
        g.params.max_iterations = g.params.channel_max[k];
  }
  printf("Creating %dx%d image, %d max iterations.\n", g.params.width, g.params.height,
    g.params.max_iterations);
  printf("Calculating image...\n");
  size_t pixel_count = (size_t) g.params.width * g.params.height;
  /* no per-thread RNG state: the GPU holds the histogram and, later, the 16-bit image */
  printf("Approximate memory needed: %.03f MiB GPU, %.03f MiB CPU\n",
    (float) (ImageBufferSize() * Channels() + pixel_count * sizeof(uint16_t)) / (1024.0 * 1024.0),
    (float) (ImageBufferSize() * Channels() + pixel_count * sizeof(uint16_t)) / (1024.0 * 1024.0));
  for (int i = 0; i < g.gpus; i++) {
    buddha_params p = g.params;
    p.device = g.params.device + i;
    int rc = buddha_create(&g.ctx[i], &p);
    if (rc != BUDDHA_OK) {
      printf("Creating the renderer on GPU %d failed: %s\n", p.device, buddha_last_error(NULL));
      Cleanup();
      return 1;
    }
  }
  g.host_buddhabrot = (uint32_t *) calloc(Channels(), ImageBufferSize());
  g.grayscale_image = (uint16_t *) calloc(pixel_count, sizeof(uint16_t));
  if (!g.host_buddhabrot || !g.grayscale_image) {
    printf("Failed allocating host buffers.\n");
    Cleanup();
    return 1;
  }
  LoadInProgressBuffer();
  uint64_t next_sample = RenderImage();
  SaveInProgressBuffer(next_sample);
  printf("Saving image.\n");
  if (Channels() == 1) {
    SaveImage(g.output_image);
  } else {
    for (int k = 0; k < Channels(); k++) {
      char name[4096];
      ChannelFileName(name, sizeof(name), g.output_image, k, 1);
      ToneMapChannel(k);
      SaveImage(name);
    }
  }
  if (g.color_mode >= 0) SaveColorImage();
  printf("Done! Output image saved: %s\n", g.output_image);
  Cleanup();
  return 0;
}
