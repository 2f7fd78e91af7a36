//@ts-nocheck
'use strict';
// A Worker-shaped object that speaks the reference's message protocol (src/worker.js:5-24) on top of the
// addon, so the reference's own main.js can run unmodified against the GPU engine:
//   const background_thread = new SiftWorker();  background_thread.onmessage = e => ...;
//   workerComputeGaussianScaleSpace(background_thread, {...})      // reference src/worker.js:29
import { computeGaussianScaleSpace, computeDifferenceOfGaussians, findCandidateKeypoints, refineCandidateKeypoints } from './background.js';

export const WorkerMessageTypes = {
  COMPUTE_GAUSSIAN_SCALE_SPACE: 'compute-gaussian-scale-space',
  RECEIVED_GAUSSIAN_SCALE_SPACE: 'received-gaussian-scale-space',
  COMPUTE_DIFFERENCE_OF_GAUSSIANS: 'compute-difference-of-gaussians',
  RECEIVED_DIFFERENCE_OF_GAUSSIANS: 'received-difference-of-gaussians',
  FIND_CANDIDATE_KEYPOINTS: 'find-candidate-keypoints',
  RECEIVED_CANDIDATE_KEYPOINTS: 'received-candidate-keypoints',
  REFINE_CANDIDATE_KEYPOINTS: 'refine-candidate-keypoints',
  RECEIVED_REFINED_KEYPOINTS: 'received-refined-keypoints',
};

export class SiftWorker {
  onmessage = null;
  postMessage(message) {
    const T = WorkerMessageTypes;
    let reply;
    switch (message.type) {                                            // background.js:18-49
      case T.COMPUTE_GAUSSIAN_SCALE_SPACE:
        reply = { type: T.RECEIVED_GAUSSIAN_SCALE_SPACE, scaleSpace: computeGaussianScaleSpace(message) };
        break;
      case T.COMPUTE_DIFFERENCE_OF_GAUSSIANS:
        reply = { type: T.RECEIVED_DIFFERENCE_OF_GAUSSIANS, differenceOfGaussians: computeDifferenceOfGaussians(message.scaleSpace) };
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
    queueMicrotask(() => this.onmessage && this.onmessage({ data: reply }));
  }
}
