// pm.hpp -- header-only C++ look-alikes of the OpenCV calls made by
// /root/reference/Points Matching/main.cpp, implemented over the C ABI of pm.h (libpm.so,
// B200 / sm_100a CUDA kernels; no CPU fallback).
//
//   main.cpp:43-46   BruteForceMatcher<L2<float> > matcher; matcher.match(d1, d2, matches)
//                    -> pm::BruteForceMatcher<pm::L2<float> >, pm::BFMatcher (4.x spelling)
//   main.cpp:49-69   good-match filter            -> pm::minMaxFilter (literal), pm::ratioTest
//   main.cpp:89-91   KeyPoint::convert(kps, pts, idx)          -> pm::KeyPoint::convert
//   main.cpp:95-98   findFundamentalMat(pts1, pts2, method...) -> pm::findFundamentalMat
//   main.cpp:127-132 computeCorrespondEpilines                 -> pm::computeCorrespondEpilines
//
// Error behaviour follows OpenCV: argument errors throw pm::Exception (code -215, like
// cv::Exception from CV_Assert), degenerate geometry returns an empty matrix, an empty
// query gives an empty result.  Types are layout-compatible with their cv:: namesakes
// (DMatch 16 B, Point2f 8 B, KeyPoint 28 B), so with OpenCV headers present a caller may
// reinterpret_cast vectors instead of copying.
#ifndef PM_HPP
#define PM_HPP
#include <array>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "pm.h"

namespace pm {

enum { NORM_L2 = 4, NORM_HAMMING = 6 };                               // cv::NORM_*
enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8 };   // cv::FM_* / CV_FM_*

class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct DMatch {   // == cv::DMatch
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    bool operator<(const DMatch &m) const { return distance < m.distance; }
};
static_assert(sizeof(DMatch) == sizeof(pm_dmatch), "DMatch must match cv::DMatch / pm_dmatch");

struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Vec3f { float val[3]; float operator[](int i) const { return val[i]; } };

struct KeyPoint {   // == cv::KeyPoint (pt, size, angle, response, octave, class_id)
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    // KeyPoint::convert(keypoints, points2f, keypointIndexes) -- main.cpp:90-91
    static void convert(const std::vector<KeyPoint> &keypoints, std::vector<Point2f> &points2f,
                        const std::vector<int> &keypointIndexes = std::vector<int>());
};

// Row-major descriptor matrix view (what cv::Mat is at main.cpp:38-40): f32 or u8 rows.
struct Descriptors {
    const void *data; int rows, cols; bool is_u8;
    Descriptors() : data(nullptr), rows(0), cols(0), is_u8(false) {}
    Descriptors(const float *p, int r, int c) : data(p), rows(r), cols(c), is_u8(false) {}
    Descriptors(const uint8_t *p, int r, int c) : data(p), rows(r), cols(c), is_u8(true) {}
    bool empty() const { return rows == 0; }
};

// The matrix findFundamentalMat returns: empty, 3x3, or -- for exactly seven points -- the 7-point solver's 1..3 real
// roots stacked as a 3k x 3 matrix (cv::Mat 9x3 when k = 3).  Row-major doubles; operator()(r, c) indexes the stack.
struct Matx33d {
    double val[27]; int n_models;
    Matx33d() : n_models(0) { std::memset(val, 0, sizeof(val)); }
    bool empty() const { return n_models == 0; }
    int rows() const { return 3 * n_models; }
    int cols() const { return 3; }
    double operator()(int r, int c) const { return val[3 * r + c]; }
};

// RAII pm_ctx.  One per host thread (a ctx is not thread-safe).
class Context {
public:
    explicit Context(int device = 0) : h_(nullptr)
    {
        const int st = pm_create(&h_, device);
        if (st != PM_OK) throw Exception(st, "pm_create failed: no sm_100 CUDA device (libpm has no CPU fallback)");
    }
    ~Context() { if (h_) pm_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    pm_ctx *handle() const { return h_; }
    // PM_OK passes, PM_EMPTY passes when allowed, anything else throws like cv::error
    int check(int st, bool allow_empty = false) const
    {
        if (st == PM_OK || (allow_empty && st == PM_EMPTY)) return st;
        throw Exception(st == PM_BAD_ARG ? -215 : st, std::string("libpm: ") + pm_last_error(h_));
    }
private:
    pm_ctx *h_;
};

inline Context &defaultContext()
{
    static thread_local Context ctx(0);
    return ctx;
}

// ---------------------------------------------------------------------------------------
// matching
// ---------------------------------------------------------------------------------------
class BFMatcher {
public:
    explicit BFMatcher(int normType = NORM_L2, bool crossCheck = false, Context *ctx = nullptr)
        : norm_(normType), cross_(crossCheck), ctx_(ctx)
    {
        if (normType != NORM_L2 && normType != NORM_HAMMING)
            throw Exception(-215, "BFMatcher: normType must be NORM_L2 or NORM_HAMMING");
    }
    // [nq][2] kNN-2 array; absent neighbours have trainIdx = -1
    void knn2(const Descriptors &q, const Descriptors &t, std::vector<DMatch> &out) const
    {
        Context &c = ctx();
        checkTypes(q, t);
        out.assign((size_t)q.rows * 2, DMatch());
        if (q.rows == 0) return;
        pm_dmatch *o = reinterpret_cast<pm_dmatch *>(out.data());
        if (norm_ == NORM_HAMMING)
            c.check(pm_knn2_hamming(c.handle(), (const uint8_t *)q.data, q.rows, (const uint8_t *)t.data, t.rows, q.cols, o));
        else if (q.is_u8)
            c.check(pm_knn2_l2_u8(c.handle(), (const uint8_t *)q.data, q.rows, (const uint8_t *)t.data, t.rows, q.cols, o));
        else
            c.check(pm_knn2_l2_f32(c.handle(), (const float *)q.data, q.rows, (const float *)t.data, t.rows, q.cols, o));
    }
    // DescriptorMatcher::match (main.cpp:46): best match per query in queryIdx order; with
    // crossCheck only mutual nearest neighbours survive
    void match(const Descriptors &q, const Descriptors &t, std::vector<DMatch> &matches) const
    {
        matches.clear();
        if (cross_) {
            Context &c = ctx();
            checkTypes(q, t);
            if (q.rows == 0 || t.rows == 0) return;
            matches.resize((size_t)q.rows);
            int n = 0;
            pm_dmatch *o = reinterpret_cast<pm_dmatch *>(matches.data());
            if (norm_ == NORM_HAMMING)
                c.check(pm_match_cross_hamming(c.handle(), (const uint8_t *)q.data, q.rows, (const uint8_t *)t.data, t.rows, q.cols, o, &n));
            else {
                if (q.is_u8) throw Exception(-215, "crossCheck L2 takes CV_32F descriptors");
                c.check(pm_match_cross_l2_f32(c.handle(), (const float *)q.data, q.rows, (const float *)t.data, t.rows, q.cols, o, &n));
            }
            matches.resize((size_t)n);
            return;
        }
        std::vector<DMatch> knn;
        knn2(q, t, knn);
        for (int i = 0; i < q.rows; ++i)
            if (knn[2 * (size_t)i].trainIdx >= 0) matches.push_back(knn[2 * (size_t)i]);
    }
    // DescriptorMatcher::knnMatch, k in {1, 2}; rows are shorter than k when the train set is
    void knnMatch(const Descriptors &q, const Descriptors &t, std::vector<std::vector<DMatch> > &matches, int k,
                  bool compactResult = false) const
    {
        if (k != 1 && k != 2) throw Exception(-215, "knnMatch: k must be 1 or 2");
        if (cross_ && k != 1) throw Exception(-215, "K == 1 && update == 0 (crossCheck needs k == 1)");
        matches.clear();
        if (cross_) {
            std::vector<DMatch> m;
            match(q, t, m);
            if (!compactResult) matches.resize((size_t)q.rows);
            for (const DMatch &d : m) {
                if (compactResult) matches.push_back(std::vector<DMatch>(1, d));
                else matches[(size_t)d.queryIdx].push_back(d);
            }
            return;
        }
        std::vector<DMatch> knn;
        knn2(q, t, knn);
        for (int i = 0; i < q.rows; ++i) {
            std::vector<DMatch> row;
            for (int j = 0; j < k; ++j)
                if (knn[2 * (size_t)i + j].trainIdx >= 0) row.push_back(knn[2 * (size_t)i + j]);
            if (!row.empty() || !compactResult) matches.push_back(row);
        }
    }
    Context &ctx() const { return ctx_ ? *ctx_ : defaultContext(); }
private:
    void checkTypes(const Descriptors &q, const Descriptors &t) const
    {
        if (q.rows < 0 || t.rows < 0 || q.cols <= 0) throw Exception(-215, "bad descriptor matrix");
        if (t.rows > 0 && (t.cols != q.cols || t.is_u8 != q.is_u8))
            throw Exception(-215, "_queryDescriptors.type() == trainDescType");
        if (norm_ == NORM_HAMMING && !q.is_u8) throw Exception(-215, "NORM_HAMMING takes CV_8U descriptors");
    }
    int norm_; bool cross_; Context *ctx_;
};

// legacy spelling used by the reference: BruteForceMatcher<L2<float> > matcher;
template <typename T> struct L2 { enum { normType = NORM_L2 }; typedef T ValueType; };
struct Hamming { enum { normType = NORM_HAMMING }; typedef unsigned char ValueType; };
template <class Distance> class BruteForceMatcher : public BFMatcher {
public:
    BruteForceMatcher() : BFMatcher((int)Distance::normType, false) {}
};

// Lowe ratio test over a [nq][2] kNN array (north_star's form of main.cpp:49-69)
inline void ratioTest(const std::vector<DMatch> &knn2, float ratio, std::vector<DMatch> &good, Context *ctx = nullptr)
{
    Context &c = ctx ? *ctx : defaultContext();
    const int nq = (int)(knn2.size() / 2);
    good.assign((size_t)nq, DMatch());
    int n = 0;
    c.check(pm_ratio_filter(c.handle(), reinterpret_cast<const pm_dmatch *>(knn2.data()), nq, ratio,
                            reinterpret_cast<pm_dmatch *>(good.data()), &n));
    good.resize((size_t)n);
}

// The reference's literal rule (main.cpp:49-69): minMatch starts at 1, maxMatch at 0,
// keep distance < min + (max - min) / 2
inline void minMaxFilter(const std::vector<DMatch> &matches, std::vector<DMatch> &good, double *minMatch = nullptr,
                         double *maxMatch = nullptr, Context *ctx = nullptr)
{
    Context &c = ctx ? *ctx : defaultContext();
    const int n_in = (int)matches.size();
    good.assign((size_t)(n_in ? n_in : 1), DMatch());
    int n = 0;
    c.check(pm_minmax_filter(c.handle(), reinterpret_cast<const pm_dmatch *>(matches.data()), n_in, 1,
                             reinterpret_cast<pm_dmatch *>(good.data()), &n, minMatch, maxMatch));
    good.resize((size_t)n);
}

inline void KeyPoint::convert(const std::vector<KeyPoint> &keypoints, std::vector<Point2f> &points2f,
                              const std::vector<int> &keypointIndexes)
{
    Context &c = defaultContext();
    std::vector<Point2f> xy(keypoints.size());
    for (size_t i = 0; i < keypoints.size(); ++i) xy[i] = keypoints[i].pt;
    if (keypointIndexes.empty()) { points2f = xy; return; }
    for (int i : keypointIndexes)
        if (i < 0 || (size_t)i >= keypoints.size())
            throw Exception(-215, "keypointIndexes has element < 0 or >= keypoints.size()");   // OpenCV: CV_Error
    points2f.assign(keypointIndexes.size(), Point2f());
    c.check(pm_gather_points(c.handle(), reinterpret_cast<const float *>(xy.data()), (int)xy.size(), keypointIndexes.data(),
                             (int)keypointIndexes.size(), reinterpret_cast<float *>(points2f.data())));
}

// ---------------------------------------------------------------------------------------
// cv::findFundamentalMat (main.cpp:95-98) over pm_find_fundamental_mat, which carries OpenCV's dispatch table:
//   N < 7 -> empty.  N == 7 (any method) -> the 7-point roots, stacked (1..3 models), mask of ones.  FM_8POINT ->
//   N-point normalised 8-point, mask of ones.  FM_RANSAC with N >= 15 -> RANSAC with the adaptive stop
//   niters = log(1 - conf) / log(1 - w^m), batches generated, solved and scored on the GPU.  Everything else
//   (FM_7POINT with N > 7 = the reference's literal call, FM_LMEDS, FM_RANSAC with N < 15) -> LMedS over 7-point
//   samples.  param1 <= 0 -> 3, param2 outside (0, 1) -> 0.99.
// The default options are OpenCV's estimator (7-point samples, symmetric-epipolar error, no refit: the mask obeys
// cv's rule err <= thr^2 exactly); FundamentalOptions::northStar() is BASELINE.json's variant.
// ---------------------------------------------------------------------------------------
struct FundamentalOptions {
    int sampleSize = 7;               // RANSAC minimal sample: 7 (OpenCV) or 8
    int metric = PM_METRIC_SYMEPI;    // OpenCV's computeError; PM_METRIC_SAMPSON for the north_star variant
    bool refit = false;               // 8-point refit on the winner's inliers (OpenCV does not refit)
    int maxIters = 1000, batch = 1024;
    uint64_t seed = 0;
    static FundamentalOptions northStar()
    {
        FundamentalOptions o; o.sampleSize = 8; o.metric = PM_METRIC_SAMPSON; o.refit = true; return o;
    }
};

inline Matx33d findFundamentalMat(const std::vector<Point2f> &points1, const std::vector<Point2f> &points2,
                                  int method = FM_RANSAC, double param1 = 3., double param2 = 0.99,
                                  std::vector<unsigned char> *mask = nullptr,
                                  const FundamentalOptions &opt = FundamentalOptions(), Context *ctx = nullptr)
{
    Context &c = ctx ? *ctx : defaultContext();
    if (points1.size() != points2.size()) throw Exception(-215, "points1/points2 count mismatch");
    if (method != FM_7POINT && method != FM_8POINT && method != FM_LMEDS && method != FM_RANSAC)
        throw Exception(-215, "findFundamentalMat: unknown method");
    const int n = (int)points1.size();
    Matx33d F;
    pm_fm_options o;
    o.sample_size = opt.sampleSize; o.metric = opt.metric; o.refit = opt.refit ? 1 : 0; o.batch = opt.batch; o.seed = opt.seed;
    std::vector<unsigned char> m((size_t)(n > 0 ? n : 1));
    int k = 0;
    const int st = c.check(pm_find_fundamental_mat(c.handle(), reinterpret_cast<const float *>(points1.data()),
                                                   reinterpret_cast<const float *>(points2.data()), n, method, param1, param2,
                                                   opt.maxIters, &o, F.val, &k, m.data()), true);
    if (st != PM_OK) return Matx33d();
    F.n_models = k;
    if (mask) { m.resize((size_t)n); *mask = m; }
    return F;
}

// cv::computeCorrespondEpilines (main.cpp:128-132): l = F x (whichImage 1) or F^T x (2), a^2 + b^2 = 1
inline void computeCorrespondEpilines(const std::vector<Point2f> &points, int whichImage, const Matx33d &F,
                                      std::vector<Vec3f> &lines, Context *ctx = nullptr)
{
    Context &c = ctx ? *ctx : defaultContext();
    if (F.empty() || (whichImage != 1 && whichImage != 2)) throw Exception(-215, "computeCorrespondEpilines: bad F / whichImage");
    lines.assign(points.size(), Vec3f());
    c.check(pm_epilines(c.handle(), reinterpret_cast<const float *>(points.data()), (int)points.size(), whichImage, F.val,
                        reinterpret_cast<float *>(lines.data())));
}

// Residuals in the correct convention x2^T F x1 (main.cpp:110-117 evaluates x1^T F x2).
inline double epipolarResiduals(const std::vector<Point2f> &points1, const std::vector<Point2f> &points2, const Matx33d &F,
                                std::vector<float> &out, int metric = PM_METRIC_SAMPSON, Context *ctx = nullptr)
{
    Context &c = ctx ? *ctx : defaultContext();
    if (F.empty() || points1.size() != points2.size()) throw Exception(-215, "epipolarResiduals: bad arguments");
    out.assign(points1.size(), 0.f);
    double mean = 0;
    c.check(pm_residuals(c.handle(), reinterpret_cast<const float *>(points1.data()), reinterpret_cast<const float *>(points2.data()),
                         (int)points1.size(), F.val, metric, out.data(), &mean));
    return mean;
}

}  // namespace pm
#endif  // PM_HPP
