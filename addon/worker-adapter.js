//@ts-nocheck
'use strict';
// A Worker-shaped object that speaks the reference's message protocol (src/worker.js:5-24) on top of the
// addon, so the reference's own main.js can run unmodified against the GPU engine:
//   const background_thread = new SiftWorker();  background_thread.onmessage = e => ...;
//   workerComputeGaussianScaleSpace(background_thread, {...})      // reference src/worker.js:29
import { computeGaussianScaleSpace, computeDifferenceOfGaussians, findCandidateKeypoints, refineCandidateKeypoints, levelPreviews } from './background.js';
import { LEVEL } from './native.js';

export const WorkerMessageTypes = {
  COMPUTE_GAUSSIAN_SCALE_SPACE: 'compute-gaussian-scale-space',
  RECEIVED_GAUSSIAN_SCALE_SPACE: 'received-gaussian-scale-space',
  RECEIVED_GAUSSIAN_BLURRED_IMAGE: 'received-gaussian-blurred-image',
  COMPUTE_DIFFERENCE_OF_GAUSSIANS: 'compute-difference-of-gaussians',
  RECEIVED_DIFFERENCE_OF_GAUSSIANS: 'received-difference-of-gaussians',
  RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE: 'received-difference-of-gaussian-image',
  FIND_CANDIDATE_KEYPOINTS: 'find-candidate-keypoints',
  RECEIVED_CANDIDATE_KEYPOINTS: 'received-candidate-keypoints',
  REFINE_CANDIDATE_KEYPOINTS: 'refine-candidate-keypoints',
  RECEIVED_REFINED_KEYPOINTS: 'received-refined-keypoints',
};

/** `new SiftWorker()` / `SiftWorker()`: an object with the two members the reference uses of its Worker --
 *  `postMessage(message)` and an assignable `onmessage` (main.js:46-47, src/worker.js:29-98).
 *  previews: also post the per-level images the reference's canvas consumes (main.js:150-165), before the stage
 *  reply as the reference does.  The per-chunk repaint messages (RECEIVED_*_CHUNK) and the candidate markers are
 *  progressive-display only and are not emitted. */
export function SiftWorker({ previews = false } = {}) {
  const T = WorkerMessageTypes;
  const worker = {
    onmessage: null,
    previews,
    postMessage: (message) => {
      let reply;
      const extra = [];
      switch (message.type) {                                          // background.js:18-49
        case T.COMPUTE_GAUSSIAN_SCALE_SPACE:
          reply = { type: T.RECEIVED_GAUSSIAN_SCALE_SPACE, scaleSpace: computeGaussianScaleSpace(message) };
          if (worker.previews) for (const p of levelPreviews(LEVEL.GAUSSIAN)) extra.push({ type: T.RECEIVED_GAUSSIAN_BLURRED_IMAGE, ...p });
          break;
        case T.COMPUTE_DIFFERENCE_OF_GAUSSIANS:
          reply = { type: T.RECEIVED_DIFFERENCE_OF_GAUSSIANS, differenceOfGaussians: computeDifferenceOfGaussians(message.scaleSpace) };
          if (worker.previews) for (const p of levelPreviews(LEVEL.DOG)) extra.push({ type: T.RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE, ...p });
          break;
        case T.FIND_CANDIDATE_KEYPOINTS:
          reply = { type: T.RECEIVED_CANDIDATE_KEYPOINTS, candidateKeypoints: findCandidateKeypoints(message) };
          break;
        case T.REFINE_CANDIDATE_KEYPOINTS:
          reply = { type: T.RECEIVED_REFINED_KEYPOINTS, refinedKeypoints: refineCandidateKeypoints(message) };
          break;
        default:
          return;
      }
      queueMicrotask(() => { if (worker.onmessage) for (const m of [...extra, reply]) worker.onmessage({ data: m }); });
    },
  };
  return worker;
}
