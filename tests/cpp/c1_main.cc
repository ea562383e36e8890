// Config C1 through the drop-in C++ class: two 640x480 frames, ORB_SLAM2::ORBextractor(1000, 1.2, 8, 15, 5)::operator() on each
// exactly the way Frame::ExtractORB calls it (src/Frame.cc:413-419), `reps` times; prints the median wall time of the two calls
// and dumps keypoints + descriptors of both frames (tools/c1_cpp.py adds ORBmatcher::SearchForInitialization through the
// drop-in ORBmatcher class and checks everything against the oracle).
//   usage: c1_main rows cols nfeatures nlevels a.raw b.raw out.bin reps
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ORBextractor.h"

static bool load(const char* path, cv::Mat& im) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    const size_t n = (size_t)im.rows * im.cols;
    const bool ok = std::fread(im.ptr(0), 1, n, f) == n;
    std::fclose(f);
    return ok;
}

int main(int argc, char** argv) {
    if (argc != 9) return 2;
    const int rows = std::atoi(argv[1]), cols = std::atoi(argv[2]), nf = std::atoi(argv[3]), nl = std::atoi(argv[4]), reps = std::atoi(argv[8]);
    cv::Mat a(rows, cols, CV_8UC1), b(rows, cols, CV_8UC1);
    if (!load(argv[5], a) || !load(argv[6], b)) return 3;
    ORB_SLAM2::ORBextractor ex(nf, 1.2f, nl, 15, 5);
    std::vector<cv::KeyPoint> ka, kb;
    cv::Mat da, db;
    std::vector<double> us;
    for (int r = 0; r < reps + 3; ++r) {
        const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ex(a, cv::Mat(), ka, da);
        ex(b, cv::Mat(), kb, db);
        const double t = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        if (r >= 3) us.push_back(t);
    }
    std::sort(us.begin(), us.end());
    std::printf("%.3f\n", us[us.size() / 2]);          // median microseconds for the two operator() calls
    FILE* f = std::fopen(argv[7], "wb");
    if (!f) return 4;
    for (int k = 0; k < 2; ++k) {
        const std::vector<cv::KeyPoint>& kp = k ? kb : ka;
        const cv::Mat& d = k ? db : da;
        int n = (int)kp.size();
        std::fwrite(&n, 4, 1, f);
        std::fwrite(kp.data(), sizeof(cv::KeyPoint), n, f);
        for (int i = 0; i < n; ++i) std::fwrite(d.ptr(i), 1, 32, f);
    }
    std::fclose(f);
    return 0;
}
