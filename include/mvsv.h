/*
 * mvsv.h -- C ABI of the B200-native stereo disparity engine (libmvsv.so).
 *
 * This is the drop-in boundary for mvStereoVision3's hot path.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repository root).  Plain
 * pointers and sizes only; no C++/torch/OpenCV types cross this line.  All functions return
 * MVSV_OK (0) or a negative error code and never throw; mvsv_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device mvsv_init fails with MVSV_ERR_CUDA.
 *
 * Threading contract (reference: one worker std::thread calls Disparity::sgbm,
 * trgt/demo.cpp:68-77,195): a ctx is thread-compatible -- any thread may call, one at a time.
 * Several ctxs (one per GPU or stream) may run concurrently.
 */
#ifndef MVSV_H_
#define MVSV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvsv_ctx mvsv_ctx;

enum {
    MVSV_OK = 0,
    MVSV_ERR_INVALID = -1,     /* bad argument / parameter outside the bit-exact contract */
    MVSV_ERR_CUDA = -2,        /* CUDA runtime error (text in mvsv_last_error)            */
    MVSV_ERR_NOMEM = -3,
    MVSV_ERR_STATE = -4,       /* call order (e.g. compute before set_*_params)           */
    MVSV_ERR_UNSUPPORTED = -5
};

/* stage bit-mask for mvsv_compute* */
enum {
    MVSV_STAGE_RECTIFY = 1,    /* cv::remap x2 + crop      src/Stereosystem.cpp:252-256 */
    MVSV_STAGE_SGBM = 2,       /* Disparity::sgbm          src/disparity.cpp:6-10       */
    MVSV_STAGE_BM = 4,         /* Disparity::bm            src/disparity.cpp:18-22      */
    MVSV_STAGE_XYZ = 8,        /* Utility::dmap2pcl loop   src/utility.cpp:242-262      */
    MVSV_STAGE_MEANS = 16      /* Utility::calcMeanDisparity per ROI  src/utility.cpp:265-285 */
};

/* Field order of the first nine ints == struct Disparity::sgbmParameters (inc/disparity.h:17-27).
 * P1/P2 are never set by the reference (src/disparity.cpp:83-90): leave 0 to get OpenCV's 2 / 5. */
typedef struct mvsv_sgbm_params {
    int minDisp, numDisp, blockSize, disp12MaxDiff, preFilterCap, uniquenessRatio;
    int speckleWindowSize, speckleRange, disparityMode; /* 1 = MODE_HH (8 paths), else MODE_SGBM (5 paths) */
    int P1, P2;
} mvsv_sgbm_params;

/* cv::StereoBM state as driven by configs/bm.yml (PREFILTER_XSOBEL, minDisparity 0). */
typedef struct mvsv_bm_params {
    int numDisp, blockSize, preFilterCap, textureThreshold, uniquenessRatio;
} mvsv_bm_params;

typedef struct mvsv_info {
    int frame_width, frame_height;   /* raw camera frame                                        */
    int width, height;               /* rectified + cropped pair == disparity map size          */
    int max_batch;
    int sgbm_minX1, sgbm_W1, sgbm_D, sgbm_Dpad, sgbm_npaths;   /* evaluated cost domain          */
    int num_rois;
    int device;
    int sgbm_td_cluster;             /* column strips (CTAs) per frame of the fused previous-row sweep at max_batch
                                        (0: independent passes)                                */
    int last_batch;                  /* stereo pairs of the last compute == frames mvsv_download copies      */
    int sgbm_s8;                     /* the last SGBM compute kept the aggregated volume as one byte per cell
                                        (npaths * P2 <= 255: every path cost is C + e, 0 <= e <= P2)              */
    int bm_col8;                     /* the StereoBM parameters keep the column-sum volume as one byte per cell
                                        (blockSize * 2 * preFilterCap <= 255, e.g. configs/bm.yml: 21 * 4 = 84)   */
} mvsv_info;

/* Create an engine on CUDA device `device` for raw frames of frame_width x frame_height, holding up to
 * max_batch stereo pairs in flight per compute call.  Replaces the cv::Ptr<cv::StereoSGBM>/StereoBM
 * objects the drivers create (trgt/demo.cpp:190-194) and Stereosystem's rectification state. */
int mvsv_init(int device, int frame_width, int frame_height, int max_batch, mvsv_ctx** out);
void mvsv_destroy(mvsv_ctx* ctx);
const char* mvsv_last_error(const mvsv_ctx* ctx); /* ctx may be NULL: error of the last failed mvsv_init */

/* Replaces the eight StereoSGBM setters + setMode of Disparity::loadSGBMParameters
 * (src/disparity.cpp:83-95).  Parameters outside the bit-exact contract are rejected:
 * numDisp in [8,256] and a multiple of 8; blockSize^2*(2*ftzero+63)+P2 <= 32767. */
int mvsv_set_sgbm_params(mvsv_ctx* ctx, const mvsv_sgbm_params* p);
/* cv::StereoBM::create(numDisp, blockSize) + setters (trgt/disparityTest.cpp:268, configs/bm.yml). */
int mvsv_set_bm_params(mvsv_ctx* ctx, const mvsv_bm_params* p);

/* Replaces Stereosystem::initRectification's map state (src/Stereosystem.cpp:214-220): float CV_32FC1
 * maps of one camera (cam 0 = left, 1 = right), frame-sized, plus mDisplayROI.  The maps are converted once
 * to OpenCV's fixed-point (CV_16SC2 + 5-bit fractions) form on the device.  stride is in bytes. */
int mvsv_upload_rectify_maps(mvsv_ctx* ctx, int cam, const float* mapx, const float* mapy, size_t stride_bytes,
                             int roi_x, int roi_y, int roi_w, int roi_h);
/* The same state built on the device from the calibration itself: replaces the call
 * cv::initUndistortRectifyMap(K, D, R, P, size, CV_32FC1, map1, map2) of Stereosystem::initRectification
 * (src/Stereosystem.cpp:214-217) together with the upload above.  K: 3x3 camera matrix (intrinsic.yml, halved for
 * binning as src/Stereosystem.cpp:203-208 does), dist: n_dist = 0, 4 or 5 coefficients k1 k2 p1 p2 [k3]
 * (mDistCoeffs), R: 3x3 rectifying rotation and P: 3x4 projection from cv::stereoRectify (mR0/mR1, mP0/mP1),
 * all row-major doubles as the cv::Mat's hold them.  The fixed-point maps equal the ones OpenCV derives from its
 * float maps (cv2 4.13: identical for the reference's three calibrations, tests/test_rectify_maps.py). */
int mvsv_set_rectification(mvsv_ctx* ctx, int cam, const double K[9], const double* dist, int n_dist,
                           const double R[9], const double P[12], int roi_x, int roi_y, int roi_w, int roi_h);
/* Stereosystem::getRectifiedImagepair(sip, factor) (src/Stereosystem.cpp:279-315): after remap + crop the pair is
 * scaled with cv::resize(roi, dst, cv::Size(0,0), factor, factor) (INTER_LINEAR, which OpenCV replaces by its 2x2
 * area average at factor 0.5 -- the value trgt/test.cpp:213 passes).  Applies to MVSV_STAGE_RECTIFY computes; width
 * and height of mvsv_get_info become cvRound(roi * factor).  factor <= 0 or == 1 switches it off. */
int mvsv_set_resize(mvsv_ctx* ctx, double factor);
/* Stereosystem::resetRectification (src/Stereosystem.cpp:317-320): inputs are taken as already rectified. */
int mvsv_reset_rectification(mvsv_ctx* ctx);

/* Q as the CV_32F copy the drivers hold (trgt/demo.cpp:179-180), row-major 4x4. */
int mvsv_set_Q(mvsv_ctx* ctx, const float q[16]);
/* ROIs (x, y, w, h quadruples in disparity-map coordinates) whose mean disparity is wanted:
 * the 81 Subimages (src/MeanDisparityDetection.cpp:80-93) and/or the 5x5 Samplepoint windows
 * (src/SamplePointDetection.cpp:38-47), already shifted by the dMapROI offset (trgt/demo.cpp:87-113). */
/* The ROIs are coordinates of the current disparity-map geometry: a call that changes width/height
 * (mvsv_upload_rectify_maps, mvsv_set_rectification, mvsv_set_resize, mvsv_reset_rectification) drops them, and
 * MVSV_STAGE_MEANS then fails with MVSV_ERR_STATE until they are set again. */
int mvsv_set_mean_rois(mvsv_ctx* ctx, const int* xywh, int n);

/* One pass of the hot path over `batch` stereo pairs held in HOST memory (pair i at base + i*frame_stride).
 * Asynchronous when the host buffers are pinned (mvsv_host_alloc): the copies run on the ctx's copy stream and the
 * call returns at once.  Pageable buffers are accepted too, but the CUDA driver then stages them itself and the call
 * returns only when the input copies have been staged (the ctx keeps no staging buffer of its own).  With
 * MVSV_STAGE_RECTIFY the inputs are raw frames (frame size),
 * otherwise rectified pairs (width x height of mvsv_get_info).  Strides in bytes. */
int mvsv_compute(mvsv_ctx* ctx, const uint8_t* left, size_t lstride, const uint8_t* right, size_t rstride,
                 size_t frame_stride, int batch, unsigned stages);
/* Same, inputs already resident in device memory (used for the HBM-resident bench figure and pipelines
 * that produce frames on the GPU). */
int mvsv_compute_device(mvsv_ctx* ctx, const uint8_t* dleft, size_t lstride, const uint8_t* dright, size_t rstride,
                        size_t frame_stride, int batch, unsigned stages);
/* Disparity::tm(Stereopair const&, cv::Mat& output, unsigned kernelSize) (src/disparity.cpp:25-58): for every pixel
 * (i, j), i < rows-kernelSize, j < cols-kernelSize, the kernelSize x kernelSize block of the left image is matched
 * by normalised cross-correlation (cv::matchTemplate TM_CCORR_NORMED) against the right image's blocks at
 * (j + x, i), x in [0, cols-j-kernelSize); out(i, j) = (uint8) x of the first maximum (cv::minMaxLoc), 0 elsewhere.
 * Synchronous: host images in (rectified size of mvsv_get_info, pair b at base + b*frame_stride), `batch` CV_8U
 * maps out (map b at out + b*height*ostride).  The ranking is evaluated in exact integer arithmetic; OpenCV's
 * floating-point scores can order exact ties and last-bit near-ties differently (tests/test_tm.py).
 * Limits: kernelSize in [1, 31], width <= 4096. */
int mvsv_tm(mvsv_ctx* ctx, const uint8_t* left, size_t lstride, const uint8_t* right, size_t rstride,
            size_t frame_stride, int batch, unsigned kernel_size, uint8_t* out, size_t ostride);
/* Copy results of the last compute back to host memory and synchronise.  Any pointer may be NULL.  The number
 * of frames copied is that of the last compute (mvsv_info.last_batch): size the buffers for it.
 *   disp  : batch x height x (dstride bytes per row) int16, CV_16S x16 fixed point, INVALID=(minD-1)*16
 *   rectL/R: batch x height x (rstride bytes) uint8
 *   xyz   : batch x height x width x 3 float (0,0,0 where disparity <= 0)
 *   means : batch x num_rois float */
int mvsv_download(mvsv_ctx* ctx, int16_t* disp, size_t dstride, uint8_t* rectL, uint8_t* rectR, size_t rstride,
                  float* xyz, float* means);
/* Pipelining inside one engine.  With two I/O slots (default: one) consecutive compute calls alternate between two
 * sets of input / result buffers while sharing one set of cost volumes: the host->device copy of call k+1 runs on
 * the copy stream while the kernels of call k execute, and mvsv_download_age(ctx, 1, ...) fetches the results of
 * call k (age 1 = the call before the last one) on a download stream while the kernels of call k+1 execute.  The
 * kernels themselves run in submission order on the ctx stream.  A slot's results must be downloaded before the
 * call after next overwrites them.  No reference counterpart (the reference computes one pair per call on the
 * CPU, src/disparity.cpp:6-10); mvsv_download == age 0. */
int mvsv_set_io_slots(mvsv_ctx* ctx, int n);
int mvsv_download_age(mvsv_ctx* ctx, int age, int16_t* disp, size_t dstride, uint8_t* rectL, uint8_t* rectR, size_t rstride,
                      float* xyz, float* means);
/* Utility::calcMinMaxDisparity (src/utility.cpp:287-304) of the last computed maps as a GPU reduction: for every
 * frame the smallest and largest disparity value > 0, minmax[2*i], minmax[2*i+1]; (0, 0) when a map has none (the
 * reference dereferences an end iterator there).  Consumed by the PLY writer's grey ramp (src/ply.cpp:62-95). */
int mvsv_download_minmax(mvsv_ctx* ctx, int16_t* minmax);
int mvsv_sync(mvsv_ctx* ctx);
/* Software pipelining over several engines (one per batch in flight): kernels that later calls submit to `ctx`
 * start only after everything submitted to `other` so far has finished.  Host->device input copies of `ctx` are not
 * held back (they run on a separate copy stream), so they -- and `other`'s device->host download -- overlap the
 * kernels instead of both engines' kernels sharing the GPU and finishing together.  No reference counterpart: the
 * reference computes one pair per call on the CPU (src/disparity.cpp:6-10). */
int mvsv_order_after(mvsv_ctx* ctx, mvsv_ctx* other);

int mvsv_get_info(const mvsv_ctx* ctx, mvsv_info* info);
/* The CUDA stream (cudaStream_t) all work of this ctx is issued on -- for CUDA-event timing by the caller. */
void* mvsv_stream(mvsv_ctx* ctx);
/* Number of kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
unsigned long long mvsv_launch_count(const mvsv_ctx* ctx);

/* CUDA-event timing on the ctx stream (the stream the kernels are launched on).  timer_start/stop bracket a
 * region; profile_enable(1) additionally brackets every kernel launch with events, and profile_read returns the
 * accumulated milliseconds and launch counts per kernel id since the last read (n >= number of kernel ids;
 * returns that number).  mvsv_kernel_name(id) names an id ("" past the end).  While per-kernel timing is enabled the
 * engine launches its kernels one after the other; without it the SGBM cost kernel and the first row scan of
 * different chunks of a batch run side by side on two streams (a little faster, but a kernel's duration then
 * includes its neighbour's). */
int mvsv_timer_start(mvsv_ctx* ctx);
int mvsv_timer_stop(mvsv_ctx* ctx, float* ms);
int mvsv_profile_enable(mvsv_ctx* ctx, int enable);
int mvsv_profile_read(mvsv_ctx* ctx, float* ms, int* counts, int n);
const char* mvsv_kernel_name(int kid);

/* Pinned host memory for zero-staging async copies. */
int mvsv_host_alloc(void** p, size_t bytes);
int mvsv_host_free(void* p);

/* Test hooks: read an internal device buffer of the last compute into host memory.
 * which: 0 = cost volume C, 1 = aggregated S (before the final right-to-left pass), 2 = raw disparity
 * (after LR check, before median), 3 = vertical-sum volume, 4 = disparity after median (before speckle),
 * 5 = BM prefiltered left, 6 = BM prefiltered right, 7 / 8 = fixed-point rectification map of camera 0 / 1
 * (roi_h x roi_w int32 pairs: x*32, y*32 rounded).  Returns bytes written or a negative error. */
/* bit 0: keep the complete aggregated S volume (all paths) readable through mvsv_debug_read(which=1).
 * bit 1: never keep a volume as bytes (mvsv_info.sgbm_s8, mvsv_info.bm_col8).
 * bits 8..15: force the number of column strips per frame of the fused sweep (0xff = force the independent passes,
 * 0xfe = force the sweep with the usual choice of strips: small batches otherwise take the independent passes).
 * bits 16..23: force the number of frames per chunk of the overlapped cost kernel / first row scan (0 = chosen by the
 * engine: about seven chunks per batch, one chunk for small batches). */
/* Environment variables read by the library (diagnostics only): MVSV_SERIAL=1 launches every kernel of a compute
 * one after the other (as per-kernel timing does) -- for profilers, which serialise kernels anyway; MVSV_NO_DSM=1
 * sends the sweep's border records through global memory also where a thread-block cluster would be used. */
int mvsv_debug_set_flags(mvsv_ctx* ctx, unsigned flags);
long long mvsv_debug_read(mvsv_ctx* ctx, int which, void* host, size_t capacity_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MVSV_H_ */
