// mean_plain_demo -- the reference's frame-loop application (trgt/mean_plain.cpp, same skeleton as trgt/demo.cpp)
// on top of the engine, with a synthetic frame source instead of the two mvBlueFOX cameras:
//   main thread : "grab" a raw pair, notify the worker, and when a new disparity map is there run the two obstacle
//                 detectors on it (trgt/demo.cpp:215-276)
//   worker      : std::thread blocked on a condition variable, runs Disparity::sgbm (trgt/mean_plain.cpp:62-82)
// The pair is rectified on the GPU (MVSV_STAGE_RECTIFY replaces Stereosystem::getRectifiedImagepair's cv::remap x2
// + crop) and the ROI means come back with the map (MVSV_STAGE_MEANS replaces 81 + ~5000 calcMeanDisparity calls).
// Prints the two framerates the reference prints after 1000 iterations (here after --frames N).
//
// build: g++ -std=c++11 -O2 -pthread -Iinclude examples/mean_plain_demo.cpp -Lmvstereovision3_b200 -lmvsv
//            -Wl,-rpath,$PWD/mvstereovision3_b200 -o mean_plain_demo
// usage: mean_plain_demo [--frames N] [--sgbm configs/sgbm.yml]
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "mvsv_detection.hpp"
#include "mvsv_disparity.hpp"

namespace {

std::mutex disparityLock;
std::condition_variable cond_var;
std::atomic<bool> running(true), newDisparityMap(false);
bool frameReady = false;
int dmap_counter = 0;
double dmap_seconds = 0.0;

const int W = 752, H = 480;                 // mvBlueFOX frame (MONO8, reference src/Camera.cpp:36)

// synthetic random-dot stereogram with a disparity ramp (same idea as mvstereovision3_b200/synth.py)
void make_pair(std::vector<unsigned char>& l, std::vector<unsigned char>& r, unsigned seed, int maxD)
{
    std::vector<unsigned char> tex((size_t)H * (W + maxD + 8));
    unsigned s = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < tex.size(); ++i) { s = s * 1664525u + 1013904223u; tex[i] = (unsigned char)(s >> 24); }
    l.resize((size_t)W * H); r.resize((size_t)W * H);
    const int TW = W + maxD + 8;
    for (int y = 0; y < H; ++y) {
        const int d = 2 + (y * (maxD - 6)) / (H - 1);
        for (int x = 0; x < W; ++x) {
            l[(size_t)y * W + x] = tex[(size_t)y * TW + x];
            r[(size_t)y * W + x] = tex[(size_t)y * TW + x + d];
        }
    }
}

struct Worker {
    mvsv_ctx* ctx; const unsigned char* left; const unsigned char* right;
    std::vector<int16_t>* dMapRaw; std::vector<float>* means; int width;
};

// trgt/mean_plain.cpp:62-82 (disparityCalcSGBM), with the rectification and the ROI means folded into the call
void disparityCalcSGBM(Worker w)
{
    while (running) {
        const auto start = std::chrono::steady_clock::now();
        std::unique_lock<std::mutex> ul(disparityLock);
        cond_var.wait(ul, [] { return frameReady || !running; });
        if (!running) break;
        frameReady = false;
        if (mvsv_compute(w.ctx, w.left, W, w.right, W, 0, 1, MVSV_STAGE_RECTIFY | MVSV_STAGE_SGBM | MVSV_STAGE_MEANS) != MVSV_OK ||
            mvsv_download(w.ctx, w.dMapRaw->data(), (size_t)w.width * 2, nullptr, nullptr, 0, nullptr, w.means->data()) != MVSV_OK) {
            std::fprintf(stderr, "compute failed: %s\n", mvsv_last_error(w.ctx));
            running = false;
            break;
        }
        newDisparityMap = true;
        dmap_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        ++dmap_counter;
    }
}

}  // namespace

int main(int argc, char** argv)
{
    int frames = 200;
    const char* yml = nullptr;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--frames")) frames = std::atoi(argv[i + 1]);
        if (!std::strcmp(argv[i], "--sgbm")) yml = argv[i + 1];
    }
    mvsv_ctx* ctx = nullptr;
    if (mvsv_init(0, W, H, 1, &ctx) != MVSV_OK) { std::fprintf(stderr, "%s\n", mvsv_last_error(nullptr)); return 1; }

    // identity-like rectification maps with a small shear, display ROI 752x479 as parameters/baseline_small gives
    std::vector<float> mx((size_t)W * H), my((size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) { mx[(size_t)y * W + x] = x + 0.25f; my[(size_t)y * W + x] = y + 0.002f * (x - W / 2); }
    for (int cam = 0; cam < 2; ++cam)
        if (mvsv_upload_rectify_maps(ctx, cam, mx.data(), my.data(), W * sizeof(float), 0, 0, W, H - 1) != MVSV_OK) return 2;
    mvsv_info info;
    mvsv_get_info(ctx, &info);

    // parameters as Disparity::loadSGBMParameters wires them (src/disparity.cpp:83-95); default = configs/sgbm.yml
    mvsv_sgbm_params p;
    std::memset(&p, 0, sizeof p);
    p.minDisp = 1; p.numDisp = 128; p.blockSize = 13; p.speckleWindowSize = 150; p.speckleRange = 2;
    if (yml) {
        std::map<std::string, double> fs;
        if (!mvsv::readFlatYaml(yml, fs)) { std::fprintf(stderr, "Unable to open disparity parameters\n"); return 3; }
        p.minDisp = (int)fs["minDisp"]; p.numDisp = (int)fs["numDisp"]; p.blockSize = (int)fs["blockSize"];
        p.disp12MaxDiff = (int)fs["disp12MaxDiff"]; p.preFilterCap = (int)fs["preFilterCap"];
        p.uniquenessRatio = (int)fs["uniquenessRatio"]; p.speckleWindowSize = (int)fs["speckleWindowSize"];
        p.speckleRange = (int)fs["speckleWindowRange"]; p.disparityMode = (int)fs["mode"] == 1;
    }
    if (mvsv_set_sgbm_params(ctx, &p) != MVSV_OK) { std::fprintf(stderr, "%s\n", mvsv_last_error(ctx)); return 4; }

    // createDMapROIS (trgt/demo.cpp:87-113): the detectors see dMapRaw(cols >= numDisp/2)
    const int pixelShift = mvsv::dMapRoiOffset(p.numDisp, info.width), cols = info.width - pixelShift, rows = info.height;
    const float Q[16] = {1, 0, 0, -376.f, 0, 1, 0, -240.f, 0, 0, 0, 607.f, 0, 0, 1.f / 118.7f, 0};
    mvsv::MeanDisparityDetection m;
    mvsv::SamplepointDetection sd;
    m.init(cols, rows, Q, 0.1f, 1.5f, pixelShift);      // trgt/demo.cpp:206-210
    sd.init(cols, rows, Q, 0.1f, 1.5f, pixelShift);
    std::vector<int> rois = m.rois();
    rois.insert(rois.end(), sd.rois().begin(), sd.rois().end());
    const int nM = (int)m.rois().size() / 4;
    if (mvsv_set_mean_rois(ctx, rois.data(), (int)rois.size() / 4) != MVSV_OK) { std::fprintf(stderr, "%s\n", mvsv_last_error(ctx)); return 5; }

    std::vector<unsigned char> L[2], R[2];
    make_pair(L[0], R[0], 1, p.minDisp + p.numDisp);
    make_pair(L[1], R[1], 2, p.minDisp + p.numDisp);
    std::vector<unsigned char> curL = L[0], curR = R[0];
    std::vector<int16_t> dMapRaw((size_t)info.width * info.height);
    std::vector<float> means(rois.size() / 4);

    Worker w{ctx, curL.data(), curR.data(), &dMapRaw, &means, info.width};
    std::thread disparity(disparityCalcSGBM, w);

    double detect_seconds = 0.0;
    int frame = 0, detections = 0, found = 0;
    while (running && dmap_counter < frames) {
        {   // Stereosystem::getRectifiedImagepair: new raw pair, then wake the worker (trgt/demo.cpp:217-222)
            std::lock_guard<std::mutex> g(disparityLock);
            curL = L[frame & 1]; curR = R[frame & 1];
            frameReady = true;
        }
        cond_var.notify_one();
        if (newDisparityMap.exchange(false)) {
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<float> mcopy;
            { std::lock_guard<std::mutex> g(disparityLock); mcopy = means; }
            sd.build(mcopy.data() + nM, 0, 0); sd.detectObstacles();                                   // :271-272
            m.build(mcopy.data(), 0, mvsv::MeanDisparityDetection::MEAN_VALUE); m.detectObstacles();    // :275-276
            found += (int)m.getFoundObstacles().size();
            detect_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            ++detections;
        }
        ++frame;
        std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    running = false;
    cond_var.notify_all();
    disparity.join();
    if (dmap_counter == 0 || detections == 0) { std::fprintf(stderr, "no frames processed\n"); return 6; }
    long long valid = 0;
    for (size_t i = 0; i < dMapRaw.size(); ++i) valid += dMapRaw[i] > 0;
    std::printf("frames %d  map %dx%d  valid px in last map %lld  sub-image obstacles per frame %.1f\n", dmap_counter, info.width,
                info.height, valid, (double)found / detections);
    std::printf("Detection Framerate: %f\n", detections / detect_seconds);
    std::printf("Disparity Framerate: %f\n", dmap_counter / dmap_seconds);
    mvsv_destroy(ctx);
    return 0;
}
