// In-memory frame source with the reference's StereoIterator interface (include/Stereo_Iterator.h:115-122:
// hasNext / getNext(StereoFrame&) / reset), for batches of synthetic or pre-decoded stereo pairs.  The reference's
// own iterators read image files (Stereo_Iterator.cpp); datasets are not available offline, and for the batched GPU
// path (ebvo_stereo_batch) frames must be in host memory anyway.  Header-only; compiled against the reference header.
#pragma once
#include <utility>
#include <vector>
#include "Stereo_Iterator.h"     // the reference header

class SyntheticStereoIterator : public StereoIterator {
public:
    // images are CV_8UC1; timestamps default to the frame index (KITTIIterator does the same, Stereo_Iterator.cpp)
    SyntheticStereoIterator(std::vector<cv::Mat> left, std::vector<cv::Mat> right, std::vector<double> timestamps = {})
        : left_(std::move(left)), right_(std::move(right)), ts_(std::move(timestamps)), next_(0) {}
    bool hasNext() override { return next_ < left_.size() && next_ < right_.size(); }
    bool getNext(StereoFrame& frame) override
    {
        if (!hasNext()) return false;
        frame.left_image = left_[next_];
        frame.right_image = right_[next_];
        frame.timestamp = next_ < ts_.size() ? ts_[next_] : (double)next_;
        ++next_;
        return true;
    }
    void reset() override { next_ = 0; }
    size_t size() const { return left_.size() < right_.size() ? left_.size() : right_.size(); }
    // host pointers of every frame, the layout ebvo_stereo_batch takes (frames must be continuous, stride = cols)
    void batch_pointers(std::vector<const unsigned char*>& L, std::vector<const unsigned char*>& R) const
    {
        L.clear(); R.clear();
        for (size_t k = 0; k < size(); ++k) { L.push_back(left_[k].data); R.push_back(right_[k].data); }
    }
private:
    std::vector<cv::Mat> left_, right_;
    std::vector<double> ts_;
    size_t next_;
};
