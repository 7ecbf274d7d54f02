// Drives the C++ mirror (include/mvsv_disparity.hpp) the way a reference driver drives inc/disparity.h
// (trgt/demo.cpp:190-199,68-77): create matcher, loadSGBMParameters(yml), Disparity::sgbm(pair, dMapRaw, matcher).
// usage: shim_demo <sgbm.yml> <left.raw> <right.raw> <W> <H> <out.raw>
#include <cstdio>
#include <vector>
#include "mvsv_disparity.hpp"

static bool slurp(const char* fn, std::vector<unsigned char>& v, size_t n)
{
    FILE* f = std::fopen(fn, "rb");
    if (!f) return false;
    v.resize(n);
    const bool ok = std::fread(v.data(), 1, n, f) == n;
    std::fclose(f);
    return ok;
}

int main(int argc, char** argv)
{
    if (argc != 7) { std::fprintf(stderr, "usage\n"); return 2; }
    const int W = std::atoi(argv[4]), H = std::atoi(argv[5]);
    std::vector<unsigned char> l, r;
    if (!slurp(argv[2], l, (size_t)W * H) || !slurp(argv[3], r, (size_t)W * H)) { std::fprintf(stderr, "read failed\n"); return 2; }
    // place the images inside larger frames and hand ROI views on (step > cols), as getRectifiedImagepair does
    mvsv::Mat bigL(H + 6, W + 40, mvsv::MVSV_8UC1), bigR(H + 6, W + 40, mvsv::MVSV_8UC1);
    for (int y = 0; y < H; ++y) {
        std::memcpy(&bigL.at<unsigned char>(y + 3, 16), &l[(size_t)y * W], W);
        std::memcpy(&bigR.at<unsigned char>(y + 3, 16), &r[(size_t)y * W], W);
    }
    Stereopair s;
    s.mLeft = bigL.roi(16, 3, W, H);
    s.mRight = bigR.roi(16, 3, W, H);

    mvsv::Matcher disparitySGBM;               // cv::StereoSGBM::create(0,0,0,...) in the reference
    Disparity::sgbmParameters para;
    if (Disparity::loadSGBMParameters("/nonexistent/sgbm.yml", disparitySGBM, para)) return 3;   // must fail like the reference
    if (!Disparity::loadSGBMParameters(argv[1], disparitySGBM, para)) return 4;
    mvsv::Mat dMapRaw;
    Disparity::sgbm(s, dMapRaw, disparitySGBM);
    if (dMapRaw.empty()) { std::fprintf(stderr, "sgbm failed: %s\n", disparitySGBM.lastError()); return 5; }
    FILE* f = std::fopen(argv[6], "wb");
    for (int y = 0; y < H; ++y) std::fwrite(&dMapRaw.at<short>(y, 0), 2, W, f);
    std::fclose(f);
    std::printf("numDisp=%d blockSize=%d mode=%d -> %dx%d CV_16S\n", para.numDisp, para.blockSize, para.disparityMode, dMapRaw.cols, dMapRaw.rows);
    return 0;
}
