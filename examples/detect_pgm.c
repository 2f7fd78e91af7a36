/* detect_pgm.c -- the C ABI from plain C: detect keypoints in a binary PGM (P5, 8-bit) and print them.
 *
 *   make -C examples && examples/detect_pgm image.pgm [octaves [minBlurLevel]]
 *
 * One line per keypoint, in the reference's order, with the eight fields of its record
 * (background.js:619-628): octave scaleLevel localX localY absoluteSigma absoluteX absoluteY interpolatedValue.
 * There is no CPU fallback: without an sm_100 GPU sift_create fails and the program says so (exit 3). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sift_b200.h"

static unsigned char *read_pgm(const char *path, int *w, int *h)
{
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  int maxv = 0;
  unsigned char *px = NULL;
  if (fscanf(f, "P5 %d %d %d", w, h, &maxv) == 3 && maxv == 255 && *w > 0 && *h > 0 && fgetc(f) != EOF) {
    px = (unsigned char *)malloc((size_t)*w * *h);
    if (px && fread(px, 1, (size_t)*w * *h, f) != (size_t)*w * *h) { free(px); px = NULL; }
  }
  fclose(f);
  return px;
}

int main(int argc, char **argv)
{
  if (argc < 2) { fprintf(stderr, "usage: %s image.pgm [octaves [minBlurLevel]]\n", argv[0]); return 2; }
  sift_ctx *ctx = NULL;
  int rc = sift_create(0, &ctx);
  if (rc != SIFT_OK) { fprintf(stderr, "sift_create: %s (status %d)\n", sift_last_error(NULL), rc); return 3; }
  int w = 0, h = 0;
  unsigned char *px = read_pgm(argv[1], &w, &h);
  if (!px) { fprintf(stderr, "%s: not a readable 8-bit binary PGM\n", argv[1]); sift_destroy(ctx); return 2; }
  sift_params prm;
  sift_default_params(&prm);                       /* the reference's defaults (worker.js:33-37) */
  if (argc > 2) prm.numberOfOctaves = atoi(argv[2]);
  if (argc > 3) prm.minBlurLevel = atof(argv[3]);
  int cap = 1 << 16, n = 0;
  sift_keypoint *kp = (sift_keypoint *)malloc((size_t)cap * sizeof *kp);
  sift_stats st;
  rc = sift_detect(ctx, px, SIFT_U8, w, h, 0, &prm, kp, cap, &n, &st);
  if (rc == SIFT_ERR_CAPACITY) {                   /* n holds the required count: never a silent truncation */
    cap = n;
    kp = (sift_keypoint *)realloc(kp, (size_t)cap * sizeof *kp);
    rc = sift_detect(ctx, px, SIFT_U8, w, h, 0, &prm, kp, cap, &n, &st);
  }
  if (rc != SIFT_OK) { fprintf(stderr, "sift_detect: %s (status %d)\n", sift_last_error(ctx), rc); return 1; }
  fprintf(stderr, "%dx%d: %d candidates, %d keypoints, %d kernel launches, %.3f ms on the device\n", w, h, st.candidates, n,
          st.kernelLaunches, st.msDevice);
  for (int i = 0; i < n; i++)
    printf("%d %d %d %d %.17g %.17g %.17g %.17g\n", kp[i].octave, kp[i].scaleLevel, kp[i].localX, kp[i].localY,
           kp[i].absoluteSigma, kp[i].absoluteX, kp[i].absoluteY, kp[i].interpolatedValue);
  free(kp);
  free(px);
  sift_destroy(ctx);
  return 0;
}
