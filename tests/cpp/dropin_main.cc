// Test driver for the drop-in ORBextractor class: reads a raw 8-bit image, runs operator() exactly the way
// Frame::ExtractORB does (src/Frame.cc:413-419) and dumps keypoints + descriptors for comparison with the oracle.
//   usage: dropin_main rows cols nfeatures nlevels in.raw out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ORBextractor.h"

int main(int argc, char** argv) {
    if (argc != 7) return 2;
    const int rows = std::atoi(argv[1]), cols = std::atoi(argv[2]), nf = std::atoi(argv[3]), nl = std::atoi(argv[4]);
    cv::Mat im(rows, cols, CV_8UC1);
    FILE* f = std::fopen(argv[5], "rb");
    if (!f || std::fread(im.ptr(0), 1, (size_t)rows * cols, f) != (size_t)rows * cols) return 3;
    std::fclose(f);
    ORB_SLAM2::ORBextractor* ex = new ORB_SLAM2::ORBextractor(nf, 1.2f, nl, 15, 5);
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    (*ex)(im, cv::Mat(), keys, desc);
    (*ex)(im, cv::Mat(), keys, desc);            // second call reuses the workspace
    int n = (int)keys.size();
    f = std::fopen(argv[6], "wb");
    std::fwrite(&n, 4, 1, f);
    std::fwrite(keys.data(), sizeof(cv::KeyPoint), n, f);
    for (int i = 0; i < n; ++i) std::fwrite(desc.ptr(i), 1, 32, f);
    int lv = ex->GetLevels();
    std::fwrite(&lv, 4, 1, f);
    std::vector<float> sf = ex->GetScaleFactors(), isig = ex->GetInverseScaleSigmaSquares();
    std::fwrite(sf.data(), 4, lv, f);
    std::fwrite(isig.data(), 4, lv, f);
    // every level of mvImagePyramid WITH the 19-px frame around the view (the reference's views sit inside padded storage too)
    for (int l = 0; l < lv; ++l) {
        const cv::Mat& L = ex->mvImagePyramid[l];
        int tr = L.rows, tc = L.cols;
        std::fwrite(&tr, 4, 1, f);
        std::fwrite(&tc, 4, 1, f);
        for (int y = -19; y < tr + 19; ++y) std::fwrite(L.ptr(0) + (long)y * (long)L.step - 19, 1, tc + 38, f);
    }
    std::fclose(f);
    delete ex;
    return 0;
}
