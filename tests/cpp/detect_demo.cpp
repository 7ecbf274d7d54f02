// Drives include/mvsv_detection.hpp the way trgt/demo.cpp:206-210,271-276 drives the reference detectors.
// usage: detect_demo <cols> <rows> <x_offset> <min_dist> <max_dist> <q.bin(16 f32)> <means.bin> <minDisp> <maxDisp>
//   means.bin = 81 sub-image means followed by the sample-point means (order of rois()).
// prints a deterministic text dump that tests/test_detection.py compares with its own restatement.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include "mvsv_detection.hpp"

static std::vector<float> slurp(const char* fn)
{
    std::vector<float> v;
    FILE* f = std::fopen(fn, "rb");
    if (!f) return v;
    float x;
    while (std::fread(&x, 4, 1, f) == 1) v.push_back(x);
    std::fclose(f);
    return v;
}

int main(int argc, char** argv)
{
    if (argc != 10) return 2;
    const int cols = std::atoi(argv[1]), rows = std::atoi(argv[2]), xoff = std::atoi(argv[3]);
    const float mind = (float)std::atof(argv[4]), maxd = (float)std::atof(argv[5]);
    std::vector<float> Q = slurp(argv[6]), means = slurp(argv[7]);
    if (Q.size() != 16) return 3;
    mvsv::MeanDisparityDetection m;
    mvsv::SamplepointDetection sd;
    m.init(cols, rows, Q.data(), mind, maxd, xoff);
    sd.init(cols, rows, Q.data(), mind, maxd, xoff);
    if (means.size() != m.rois().size() / 4 + sd.rois().size() / 4) { std::fprintf(stderr, "means size\n"); return 4; }
    std::printf("range %.9g %.9g\n", m.getRangeDisparity().first, m.getRangeDisparity().second);
    std::printf("nroi %zu %zu\n", m.rois().size() / 4, sd.rois().size() / 4);
    std::printf("roi0 %d %d %d %d | sp0 %d %d %d %d\n", m.rois()[0], m.rois()[1], m.rois()[2], m.rois()[3], sd.rois()[0],
                sd.rois()[1], sd.rois()[2], sd.rois()[3]);
    m.build(means.data(), 0, mvsv::MeanDisparityDetection::MEAN_DISTANCE);   // falls through into MEAN_VALUE
    std::printf("mode %d ndist %zu\n", m.getDetectionMode(), m.getMeanDistanceMap().size());
    for (size_t i = 0; i < m.getMeanDistanceMap().size(); i += 9) std::printf("dist %zu %.9g\n", i, m.getMeanDistanceMap()[i]);
    m.detectObstacles();
    std::printf("found_mean %zu counter %d\n", m.getFoundObstacles().size(), m.getObstacleCounter());
    const std::vector<mvsv::Subimage> foundM = m.getFoundObstacles();     // getters return copies, like the reference
    for (size_t i = 0; i < m.getFoundPoints().size(); ++i) {
        const mvsv::Subimage& s = foundM[i];
        const float* p = m.getFoundPoints()[i].v;
        std::printf("M %d %d %.9g %.9g %.9g %.9g\n", s.roi_center.x, s.roi_center.y, s.value, p[0], p[1], p[2]);
    }
    sd.build(means.data() + m.rois().size() / 4, 0, 0);
    sd.detectObstacles();
    std::printf("found_sp %zu counter %d\n", sd.getFoundObstacles().size(), sd.getObstacleCounter());
    const std::vector<mvsv::Samplepoint> foundS = sd.getFoundObstacles();
    for (size_t i = 0; i < sd.getFoundPoints().size(); ++i) {
        const mvsv::Samplepoint& s = foundS[i];
        const float* p = sd.getFoundPoints()[i].v;
        std::printf("S %d %d %.9g %.9g %.9g %.9g\n", s.center.x, s.center.y, s.value, p[0], p[1], p[2]);
    }
    mvsv::ply writer("Hagen Hiller", "obstacle pointcloud", (short)std::atoi(argv[8]), (short)std::atoi(argv[9]));
    std::ostringstream os;
    writer.write(os, m.getFoundPoints(), mvsv::ply::WITH_COLOR);
    std::printf("PLY\n%s", os.str().c_str());
    return 0;
}
