//@ts-nocheck
'use strict';
// Drop-in for the reference's src/sift.js: the same five export names, argument shapes and in-place
// behaviour, each forwarding to the CUDA engine through the N-API addon (no arithmetic happens here).
//   SIFT_blurMatrix2DChunk       reference src/sift.js:72
//   SIFT_subtractMatrix2DChunk   reference src/sift.js:154
//   SIFT_findExtremas            reference src/sift.js:212
//   SIFT_generateGradientVector  reference src/sift.js:333
//   SIFT_generateHessianMatrix   reference src/sift.js:377
// Images may be Matrix2D (Array of rows: compat, slow) or {data: Float64Array, width, height} (fast).
import { native, context, toF64, toMatrix2D } from './native.js';

function writeBack(output, data, cols, { x1, y1, x2, y2 }) {
  if (Array.isArray(output)) {
    for (let y = y1; y < y2; y++) for (let x = x1; x < x2; x++) output[y][x] = data[y * cols + x];   // sift.js:137
  } else if (output.data !== data) {
    for (let y = y1; y < y2; y++) output.data.set(data.subarray(y * cols + x1, y * cols + x2), y * cols + x1);
  }
}

function chunkOf(data, cols, { x1, y1, x2, y2 }, asMatrix) {
  const w = Math.max(0, x2 - x1), h = Math.max(0, y2 - y1);
  const chunk = new Float64Array(w * h);
  for (let y = 0; y < h; y++) chunk.set(data.subarray((y1 + y) * cols + x1, (y1 + y) * cols + x2), y * w);
  return asMatrix ? toMatrix2D(chunk, h, w) : { data: chunk, width: w, height: h };
}

export function SIFT_blurMatrix2DChunk(input, output, sigma, chunk_boundary) {
  const src = toF64(input);
  const dst = Array.isArray(output) ? new Float64Array(src.rows * src.cols) : output.data;
  const { x1, y1, x2, y2 } = chunk_boundary;
  native.blurChunk(context(), src.data, src.rows, src.cols, dst, sigma, x1, y1, x2, y2);
  writeBack(output, dst, src.cols, chunk_boundary);
  return chunkOf(dst, src.cols, chunk_boundary, Array.isArray(input));
}

export function SIFT_subtractMatrix2DChunk(input_pair, output, chunk_boundary) {
  const a = toF64(input_pair[0]), b = toF64(input_pair[1]);
  const dst = Array.isArray(output) ? new Float64Array(a.rows * a.cols) : output.data;
  const { x1, y1, x2, y2 } = chunk_boundary;
  native.subtractChunk(context(), a.data, b.data, a.rows, a.cols, dst, x1, y1, x2, y2);
  writeBack(output, dst, a.cols, chunk_boundary);
  return chunkOf(dst, a.cols, chunk_boundary, Array.isArray(input_pair[0]));
}

export function SIFT_findExtremas(image_trio, scales_per_octave) {
  const d = image_trio.map(toF64);
  const r = native.findExtremas(context(), d[0].data, d[1].data, d[2].data, d[1].rows, d[1].cols,
                                scales_per_octave, 0.015, 0.8);                    // sift.js:285, 293
  const list = (n, xy, value) => Array.from({ length: n }, (_, i) => ({ x: xy[2 * i], y: xy[2 * i + 1], value: value[i] }));
  return {
    candidateKeypoints: list(r.nCand, r.candXY, r.candValue),
    lowContrastKeypoints: list(r.nLow, r.lowXY, r.lowValue),
  };
}

function gradHess(o, s, m, n, difference_of_gaussians) {
  const trio = [s - 1, s, s + 1].map(i => toF64(difference_of_gaussians[o][i].image));
  return native.gradientHessian(context(), trio[0].data, trio[1].data, trio[2].data, trio[1].rows, trio[1].cols, m, n);
}

export function SIFT_generateGradientVector(o, s, m, n, difference_of_gaussians) {
  return Array.from(gradHess(o, s, m, n, difference_of_gaussians).subarray(0, 3));
}

export function SIFT_generateHessianMatrix(o, s, m, n, difference_of_gaussians) {
  const gh = gradHess(o, s, m, n, difference_of_gaussians);
  return [Array.from(gh.subarray(3, 6)), Array.from(gh.subarray(6, 9)), Array.from(gh.subarray(9, 12))];
}
