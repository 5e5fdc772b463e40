// match_and_estimate.cpp -- host C++ over the C ABI: the matching + geometry stages of the
// reference program (/root/reference/Points Matching/main.cpp:43-98 and 127-132) with the
// OpenCV calls replaced by their pm:: look-alikes (include/pm.hpp -> libpm.so -> sm_100a
// kernels).  Detection/description and drawing are out of scope, so descriptors and keypoint
// coordinates come from raw files instead of SURF:
//
//   match_and_estimate <mode> <desc1.f32> <n1> <desc2.f32> <n2> <dim> <kp1.f32> <kp2.f32>
//     mode "literal": match() k=1 -> min/max-midpoint filter -> findFundamentalMat(FM_7POINT)
//     mode "ratio"  : knnMatch(k=2) -> ratio 0.75        -> findFundamentalMat(FM_RANSAC, 1 px)
//
// Prints one line per stage in a fixed format that tests/test_gpu_parity.py parses.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "pm.hpp"

template <typename T> static std::vector<T> read_raw(const char *path, size_t count)
{
    std::vector<T> v(count);
    std::ifstream f(path, std::ios::binary);
    if (!f || !f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(count * sizeof(T)))) {
        std::fprintf(stderr, "cannot read %zu elements from %s\n", count, path);
        std::exit(2);
    }
    return v;
}

int main(int argc, char **argv)
{
    if (argc != 9) {
        std::fprintf(stderr, "usage: %s literal|ratio desc1 n1 desc2 n2 dim kp1 kp2\n", argv[0]);
        return 2;
    }
    const bool literal = std::strcmp(argv[1], "literal") == 0;
    const int n1 = std::atoi(argv[3]), n2 = std::atoi(argv[5]), dim = std::atoi(argv[6]);
    const std::vector<float> d1 = read_raw<float>(argv[2], (size_t)n1 * dim), d2 = read_raw<float>(argv[4], (size_t)n2 * dim);
    const std::vector<float> xy1 = read_raw<float>(argv[7], (size_t)n1 * 2), xy2 = read_raw<float>(argv[8], (size_t)n2 * 2);
    std::vector<pm::KeyPoint> keyPoint1((size_t)n1), keyPoint2((size_t)n2);
    for (int i = 0; i < n1; ++i) keyPoint1[(size_t)i].pt = pm::Point2f(xy1[2 * (size_t)i], xy1[2 * (size_t)i + 1]);
    for (int i = 0; i < n2; ++i) keyPoint2[(size_t)i].pt = pm::Point2f(xy2[2 * (size_t)i], xy2[2 * (size_t)i + 1]);
    const pm::Descriptors imageDesc1(d1.data(), n1, dim), imageDesc2(d2.data(), n2, dim);

    try {
        pm::BruteForceMatcher<pm::L2<float> > matcher;            // the matcher named at main.cpp:43
        std::vector<pm::DMatch> goodMatchePoints;
        if (literal) {
            std::vector<pm::DMatch> matchePoints;
            matcher.match(imageDesc1, imageDesc2, matchePoints);  // main.cpp:46
            double minMatch = 0, maxMatch = 0;
            pm::minMaxFilter(matchePoints, goodMatchePoints, &minMatch, &maxMatch);   // main.cpp:49-69
            std::printf("matches %zu min %.9g max %.9g\n", matchePoints.size(), minMatch, maxMatch);
        } else {
            std::vector<pm::DMatch> knn;
            matcher.knn2(imageDesc1, imageDesc2, knn);
            pm::ratioTest(knn, 0.75f, goodMatchePoints);
            std::printf("matches %zu\n", knn.size() / 2);
        }
        std::printf("good %zu\n", goodMatchePoints.size());
        std::vector<int> pointIndexes1, pointIndexes2;            // main.cpp:71-79
        for (const pm::DMatch &m : goodMatchePoints) {
            std::printf("g %d %d %.9g\n", m.queryIdx, m.trainIdx, m.distance);
            pointIndexes1.push_back(m.queryIdx);
            pointIndexes2.push_back(m.trainIdx);
        }
        std::vector<pm::Point2f> selPoints1, selPoints2;          // main.cpp:89-91
        pm::KeyPoint::convert(keyPoint1, selPoints1, pointIndexes1);
        pm::KeyPoint::convert(keyPoint2, selPoints2, pointIndexes2);

        std::vector<unsigned char> mask;
        const pm::Matx33d fundemental = literal                   // main.cpp:95-98
            ? pm::findFundamentalMat(selPoints1, selPoints2, pm::FM_7POINT, 3., 0.99, &mask)
            : pm::findFundamentalMat(selPoints1, selPoints2, pm::FM_RANSAC, 1., 0.99, &mask);
        if (fundemental.empty()) { std::printf("F empty\n"); return 0; }
        std::printf("F");
        for (int i = 0; i < 9; ++i) std::printf(" %.17g", fundemental.val[i]);
        std::printf("\n");
        size_t inl = 0;
        for (unsigned char m : mask) inl += m;
        std::printf("inliers %zu\n", inl);
        std::vector<float> res;                                   // main.cpp:103-123, x2^T F x1 convention
        const double mean = pm::epipolarResiduals(selPoints1, selPoints2, fundemental, res);
        std::printf("mean_sampson %.9g\n", mean);
        std::vector<pm::Vec3f> lines1;                            // main.cpp:127-132
        pm::computeCorrespondEpilines(selPoints1, 1, fundemental, lines1);
        if (!lines1.empty()) std::printf("line0 %.9g %.9g %.9g\n", lines1[0][0], lines1[0][1], lines1[0][2]);
    } catch (const pm::Exception &e) {
        std::fprintf(stderr, "pm::Exception %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
