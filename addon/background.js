//@ts-nocheck
'use strict';
// Drop-in for the four stage functions behind the reference's worker message switch (background.js:14-50),
// as plain exported functions with the request field names of src/worker.js:40-98 and the reply schemas
// of background.js:233-236, 350-353, 446-449, 681-684 -- plus one fused detect() that keeps the pyramid
// on the GPU.  The postMessage transport is gone; addon/worker-adapter.js puts it back for callers that
// still speak WorkerMessageTypes.
import { native, context, toPixels, toF64, toMatrix2D, decodeKeypoints, decodeCandidates, encodeCandidates, LEVEL } from './native.js';

// Which replies does the device still hold?  The reference consumes the pyramid IN each request (background.js:258,
// 359, 455); here the context keeps it on the GPU, so a stage may only skip the upload when the payload is the very
// reply object this module produced AND nothing has rebuilt the context's pyramid since (native.pyramidSerial changes
// on every build / detect / setLevel).  Anything else -- another image in between, a cloned or edited payload -- is
// uploaded (DoG) or recomputed from the payload (scale space), exactly as the reference would have used it.
const resident = { gaussian: null, dog: null, serial: -1 };
const holds = (kind, payload) => payload != null && resident[kind] === payload && native.pyramidSerial(context()) === resident.serial;

const paramsOf = (o = {}) => ({
  numberOfOctaves: o.numberOfOctaves ?? 5, scalesPerOctave: o.scalesPerOctave ?? 3,         // worker.js:33-34
  minBlurLevel: o.minBlurLevel ?? 0.8, assumedBlur: o.assumedBlur ?? 0.5,                   // worker.js:35-36
  contrastThreshold: o.contrastThreshold ?? 0.015, preFilterFactor: o.preFilterFactor ?? 0.8,
  edgeRatio: o.edgeRatio ?? 10, maxIterations: o.maxIterations ?? 5, offsetBound: o.offsetBound ?? 0.6,
  minInterpixelDistance: o.minInterpixelDistance ?? 0.5,
});

function readLevels(kind, matrices) {
  const info = native.pyramidInfo(context());
  const out = [];
  for (let o = 0; o < info.octaves; o++) {
    const levels = [];
    for (let s = 0; s < info.levels - (kind === LEVEL.DOG ? 1 : 0); s++) {
      const l = native.getLevel(context(), kind, o, s);
      levels.push({ blurLevel: l.blurLevel,
                    image: matrices ? toMatrix2D(l.data, l.height, l.width) : { data: l.data, width: l.width, height: l.height } });
    }
    out.push(levels);
  }
  return out;
}

export const PREVIEW = { GRAY: 0, SIGMOID: 1, MINMAX: 2 };

/** The ImageData payloads the reference posts beside the stage replies (SURVEY 8f-3), one per level in the
 *  reference's order: Gaussian levels as grey (background.js:136-143, 212-220; image-utils.js:171-220), DoG
 *  levels min/max-normalised (background.js:331-338; matrix2d.js:169-192).  Each entry is
 *  {octave, imageData:{width, height, data:Uint8ClampedArray}} -- an ImageData-shaped plain object (wrap it
 *  with `new ImageData(data, width, height)` in a browser). */
export function levelPreviews(kind) {
  const info = native.pyramidInfo(context());
  const out = [];
  for (let o = 0; o < info.octaves; o++)
    for (let s = 0; s < info.levels - (kind === LEVEL.DOG ? 1 : 0); s++) {
      const p = native.levelPreview(context(), kind, o, s, kind === LEVEL.DOG ? PREVIEW.MINMAX : PREVIEW.GRAY, 1);
      out.push({ octave: o, imageData: { width: p.width, height: p.height, data: p.data } });
    }
  return out;
}

/** One chunk of a DoG level as the reference paints it while subtracting: sigmoid-normalised with
 *  coefficient 5 (background.js:303-317, matrix2d.js:148-156), cropped to the half-open chunk {x1,y1,x2,y2}. */
export function dogChunkPreview(octave, scale, chunk) {
  const p = native.levelPreview(context(), LEVEL.DOG, octave, scale, PREVIEW.SIGMOID, 5);
  const cw = chunk.x2 - chunk.x1, ch = chunk.y2 - chunk.y1;
  const data = new Uint8ClampedArray(cw * ch * 4);
  for (let y = 0; y < ch; y++) data.set(p.data.subarray(((chunk.y1 + y) * p.width + chunk.x1) * 4, ((chunk.y1 + y) * p.width + chunk.x2) * 4), y * cw * 4);
  return { imageData: { width: cw, height: ch, data }, dx: chunk.x1, dy: chunk.y1 };
}

/** background.js:71 -- request {inputImage, numberOfOctaves, scalesPerOctave, minBlurLevel, assumedBlur, chunkSize}. */
export function computeGaussianScaleSpace(request, { matrices = Array.isArray(request.inputImage) } = {}) {
  const px = toPixels(request.inputImage);
  native.buildScaleSpace(context(), px.data, px.width, px.height, px.dtype, paramsOf(request));
  const reply = readLevels(LEVEL.GAUSSIAN, matrices);
  resident.gaussian = reply; resident.dog = null; resident.serial = native.pyramidSerial(context());
  return reply;
}

/** background.js:258 -- the DoG of the pyramid the context holds (formed by the blur kernels). */
export function computeDifferenceOfGaussians(scaleSpace, { matrices = true } = {}) {
  if (scaleSpace == null || holds('gaussian', scaleSpace)) {       // the blur kernels already formed it, unrounded
    const reply = readLevels(LEVEL.DOG, matrices);
    if (scaleSpace != null) resident.dog = reply;
    return reply;
  }
  // a scale space this context does not hold: D[o][s-1] = S[o][s-1] - S[o][s] from the payload (background.js:269-330)
  const reply = [];
  for (const octave of scaleSpace) {
    const levels = [];
    for (let s = 1; s < octave.length; s++) {
      const a = toF64(octave[s - 1].image), b = toF64(octave[s].image);
      const out = new Float64Array(a.rows * a.cols);
      native.subtractChunk(context(), a.data, b.data, a.rows, a.cols, out, 0, 0, a.cols, a.rows);
      levels.push({ blurLevel: octave[s - 1].blurLevel,
                    image: matrices ? toMatrix2D(out, a.rows, a.cols) : { data: out, width: a.cols, height: a.rows } });
    }
    reply.push(levels);
  }
  return reply;
}

/** Make the context hold the DoG pyramid of a request (upload unless it is the resident one). */
function ensureDogResident(dog, request) {
  if (dog == null || holds('dog', dog)) return;
  const first = toF64(dog[0][0].image);
  const prm = paramsOf(request);
  prm.numberOfOctaves = dog.length; prm.scalesPerOctave = dog[0].length - 2;
  native.setPyramidShape(context(), first.cols, first.rows, prm);
  for (let o = 0; o < dog.length; o++)
    for (let s = 0; s < dog[o].length; s++) native.setLevel(context(), LEVEL.DOG, o, s, Float32Array.from(toF64(dog[o][s].image).data));
  resident.gaussian = null; resident.dog = dog; resident.serial = native.pyramidSerial(context());
}

/** background.js:359 -- request {differenceOfGaussians, octaveBaseImages (unused), scalesPerOctave}. */
export function findCandidateKeypoints(request) {
  ensureDogResident(request.differenceOfGaussians, request);
  const r = native.findCandidates(context(), null, false);
  const info = native.pyramidInfo(context());
  const reply = [];
  for (let o = 0; o < info.octaves; o++) {
    reply.push([]);
    for (let s = 1; s < info.levels - 2; s++) reply[o].push({ scaleLevel: s, localExtremas: [] });
  }
  for (const c of decodeCandidates(r.records, r.count)) reply[c.octave][c.scaleLevel - 1].localExtremas.push({ x: c.x, y: c.y, value: c.value });
  return reply;
}

/** background.js:455 -- request {differenceOfGaussians, scalesPerOctave, numberOfOctaves, candidateKeypoints,
 *  minBlurLevel, minInterpixelDistance}. */
export function refineCandidateKeypoints(request) {
  ensureDogResident(request.differenceOfGaussians, request);
  const flat = [];
  for (let octave = 0; octave < request.numberOfOctaves; octave++)                     // background.js:468-471
    for (let scale_i = 0; scale_i < request.scalesPerOctave; scale_i++) {
      const entry = request.candidateKeypoints[octave][scale_i];
      for (const e of entry.localExtremas) flat.push({ octave, scaleLevel: entry.scaleLevel, x: e.x, y: e.y, value: e.value });
    }
  const r = native.refine(context(), paramsOf(request), encodeCandidates(flat), flat.length);
  return decodeKeypoints(r.records, r.count);
}

/** The whole chain main.js:111 -> 239 -> 274 -> 325 in one device-resident call. */
export function detect(image, options = {}) {
  const px = toPixels(image);
  const r = native.detect(context(), px.data, px.width, px.height, px.dtype, paramsOf(options));
  return { keypoints: decodeKeypoints(r.records, r.count), stats: r.stats };
}

/** n equally sized frames packed back to back in one typed array; frames overlap on the GPU. */
export function detectBatch(frames, width, height, nImages, dtype, options = {}, capacity = nImages * 32768) {
  const r = native.detectBatch(context(), frames, width, height, nImages, dtype, paramsOf(options), capacity);
  const all = decodeKeypoints(r.records, r.offsets[nImages]);
  return { keypoints: Array.from({ length: nImages }, (_, i) => all.slice(r.offsets[i], r.offsets[i + 1])), stats: r.stats };
}
