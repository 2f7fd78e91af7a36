// Loads the N-API addon (sift_b200.node, built by addon/Makefile) and decodes its raw records.
// No CPU fallback: if the addon or an sm_100 GPU is missing, importing / creating throws.
import { createRequire } from 'node:module';
const require = createRequire(import.meta.url);
export const native = require('./sift_b200.node');

export const DTYPE = { U8: 0, F32: 1, F64: 2, RGBA8: 3 };
export const LEVEL = { GAUSSIAN: 0, DOG: 1 };

let defaultContext = null;
/** One context per GPU per thread, like the reference's single worker (main.js:46). */
export function context(device = 0) {
  if (defaultContext === null) defaultContext = native.create(device);
  return defaultContext;
}

/** Matrix2D (Array of rows, matrix2d.js:5-31) or {data, width, height} -> {data: Float64Array, rows, cols}. */
export function toF64(image) {
  if (Array.isArray(image)) {
    const rows = image.length, cols = image[0].length;
    const data = new Float64Array(rows * cols);
    for (let i = 0; i < rows; i++) data.set(image[i], i * cols);
    return { data, rows, cols };
  }
  const data = image.data instanceof Float64Array ? image.data : Float64Array.from(image.data);
  return { data, rows: image.height, cols: image.width };
}

/** Any accepted image -> {data, width, height, dtype} without a copy for typed arrays. */
export function toPixels(image) {
  if (Array.isArray(image)) {
    const m = toF64(image);
    return { data: m.data, width: m.cols, height: m.rows, dtype: DTYPE.F64 };
  }
  const { data, width, height } = image;
  if (data instanceof Uint8ClampedArray && data.length === width * height * 4) return { data, width, height, dtype: DTYPE.RGBA8 };
  if (data instanceof Uint8Array || data instanceof Uint8ClampedArray) return { data, width, height, dtype: DTYPE.U8 };
  if (data instanceof Float32Array) return { data, width, height, dtype: DTYPE.F32 };
  if (data instanceof Float64Array) return { data, width, height, dtype: DTYPE.F64 };
  throw new TypeError('image must be a Matrix2D or {data: TypedArray, width, height}');
}

export function toMatrix2D(data, rows, cols) {
  const out = new Array(rows);
  for (let i = 0; i < rows; i++) out[i] = Array.from(data.subarray(i * cols, (i + 1) * cols));
  return out;
}

/** 80-byte sift_keypoint records (include/sift_b200.h) -> the reference's objects (background.js:619-628)
 *  plus the interpolation offsets and the DoG value of the originating extremum. */
export function decodeKeypoints(buffer, count) {
  const v = new DataView(buffer);
  const out = new Array(count);
  for (let i = 0, o = 0; i < count; i++, o += 80) {
    out[i] = {
      octave: v.getInt32(o, true), scaleLevel: v.getInt32(o + 4, true),
      localX: v.getInt32(o + 8, true), localY: v.getInt32(o + 12, true),
      absoluteSigma: v.getFloat64(o + 16, true), absoluteX: v.getFloat64(o + 24, true),
      absoluteY: v.getFloat64(o + 32, true), interpolatedValue: v.getFloat64(o + 40, true),
      offset: [v.getFloat32(o + 48, true), v.getFloat32(o + 52, true), v.getFloat32(o + 56, true)],
      dogValue: v.getFloat32(o + 60, true),
    };
  }
  return out;
}

/** 24-byte sift_candidate records -> {octave, scaleLevel, x, y, value}. */
export function decodeCandidates(buffer, count) {
  const v = new DataView(buffer);
  const out = new Array(count);
  for (let i = 0, o = 0; i < count; i++, o += 24)
    out[i] = { octave: v.getInt32(o, true), scaleLevel: v.getInt32(o + 4, true), x: v.getInt32(o + 8, true),
               y: v.getInt32(o + 12, true), value: v.getFloat32(o + 16, true) };
  return out;
}

export function encodeCandidates(list) {
  const buffer = new ArrayBuffer(Math.max(1, list.length) * 24);
  const v = new DataView(buffer);
  list.forEach((c, i) => {
    const o = i * 24;
    v.setInt32(o, c.octave, true); v.setInt32(o + 4, c.scaleLevel, true); v.setInt32(o + 8, c.x, true);
    v.setInt32(o + 12, c.y, true); v.setFloat32(o + 16, c.value, true);
  });
  return buffer;
}
