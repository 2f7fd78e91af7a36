// Runs the UNMODIFIED reference (a copy of the reference tree under baseline/_ref/, made by __graft_entry__.build())
// under Node.js, the way SURVEY.md section 8c prescribes: the page's four requests (main.js:111 -> 239 -> 274 -> 325)
// are driven through the reference's own senders (src/worker.js:29-98) into its own onmessage switch
// (background.js:14-50).  Browser globals the worker script expects are shimmed: `onmessage` / `postMessage`
// (module code is strict, the bare assignment at background.js:14 needs the global to exist), an `OffscreenCanvas`
// whose 2d context hands out ImageData-shaped objects (image-utils.js:179-181), and a silent console.log (the
// reference logs every keypoint decision, background.js:581-672).
//
//   node baseline/run_reference.mjs <crops.bin> <width> <height> <n_crops> <octaves> <spo> <minBlur> <assumedBlur> [threads]
//
// crops.bin: n_crops grayscale u8 images of width x height back to back.  One image per worker_thread (BASELINE.md
// section 3).  Prints one JSON line: {seconds, images, keypoints:[...], threads, node}.
import { Worker, isMainThread, parentPort, workerData } from 'node:worker_threads';
import { readFileSync } from 'node:fs';
import { availableParallelism } from 'node:os';
import { fileURLToPath, pathToFileURL } from 'node:url';
import path from 'node:path';

const here = path.dirname(fileURLToPath(import.meta.url));

async function detectOne(u8, width, height, prm) {
  const outbox = [];
  globalThis.onmessage = null;
  globalThis.postMessage = (m) => { outbox.push(m); };
  globalThis.OffscreenCanvas = class {
    constructor(w, h) { this.width = w; this.height = h; }
    getContext() { return { createImageData: (w, h) => ({ width: w, height: h, data: new Uint8ClampedArray(4 * w * h) }) }; }
  };
  const log = console.log;
  console.log = () => {};
  const ref = pathToFileURL(path.join(here, '_ref')).href;
  await import(ref + '/background.js');                       // installs globalThis.onmessage
  const send = await import(ref + '/src/worker.js');
  const T = send.WorkerMessageTypes;
  const handle = { postMessage: (m) => globalThis.onmessage({ data: m }) };
  const reply = (type) => {
    const hit = outbox.filter((m) => m && m.type === type);
    outbox.length = 0;
    if (hit.length !== 1) throw new Error('expected one ' + type + ', got ' + hit.length);
    return hit[0];
  };
  // Matrix2D = Array of rows of Numbers in [0, 1] (image-utils.js:114: v / 255)
  const image = new Array(height);
  for (let y = 0; y < height; y++) {
    const row = new Array(width);
    for (let x = 0; x < width; x++) row[x] = u8[y * width + x] / 255.0;
    image[y] = row;
  }
  const t0 = process.hrtime.bigint();
  send.workerComputeGaussianScaleSpace(handle, { input_image: image, number_of_octaves: prm.octaves, scales_per_octave: prm.spo,
                                                 min_blur_level: prm.minBlur, assumed_blur: prm.assumedBlur, chunk_size: 32 });
  const ss = reply(T.RECEIVED_GAUSSIAN_SCALE_SPACE).scaleSpace;
  send.workerComputeDifferenceOfGaussians(handle, ss);
  const dog = reply(T.RECEIVED_DIFFERENCE_OF_GAUSSIANS).differenceOfGaussians;
  send.workerFindCandidateKeypoints(handle, dog, ss.map((o) => o[0].image), prm.spo);
  const cands = reply(T.RECEIVED_CANDIDATE_KEYPOINTS).candidateKeypoints;
  send.workerRefineCandidateKeypoints(handle, dog, cands, prm.spo, prm.octaves, prm.minBlur);
  const kps = reply(T.RECEIVED_REFINED_KEYPOINTS).refinedKeypoints;
  const seconds = Number(process.hrtime.bigint() - t0) / 1e9;
  console.log = log;
  return { seconds, keypoints: kps.length };
}

if (isMainThread) {
  const [file, w, h, n, octaves, spo, minBlur, assumedBlur, threadsArg] = process.argv.slice(2);
  const width = +w, height = +h, nCrops = +n;
  const bytes = readFileSync(file);
  const threads = Math.max(1, Math.min(+threadsArg || availableParallelism(), nCrops));
  const prm = { octaves: +octaves, spo: +spo, minBlur: +minBlur, assumedBlur: +assumedBlur };
  const results = new Array(nCrops);
  let next = 0, running = 0;
  const t0 = process.hrtime.bigint();
  await new Promise((resolve, reject) => {
    const launch = () => {
      while (running < threads && next < nCrops) {
        const i = next++;
        running++;
        const wk = new Worker(fileURLToPath(import.meta.url), {
          workerData: { u8: bytes.subarray(i * width * height, (i + 1) * width * height), width, height, prm },
          resourceLimits: { maxOldGenerationSizeMb: 8192 },
        });
        wk.on('message', (r) => { results[i] = r; });
        wk.on('error', reject);
        wk.on('exit', () => { running--; if (next >= nCrops && running === 0) resolve(); else launch(); });
      }
    };
    launch();
  });
  const seconds = Number(process.hrtime.bigint() - t0) / 1e9;
  process.stdout.write(JSON.stringify({ seconds, images: nCrops, threads, node: process.version,
                                        keypoints: results.map((r) => r.keypoints), perImageSeconds: results.map((r) => r.seconds) }) + '\n');
} else {
  const { u8, width, height, prm } = workerData;
  detectOne(u8, width, height, prm).then((r) => parentPort.postMessage(r));
}
