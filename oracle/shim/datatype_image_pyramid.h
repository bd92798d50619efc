// ORACLE SHIM (test infrastructure).  Stand-in for Slam_Utility's `datatype_image_pyramid.h`; semantics
// frozen in SURVEY.md Appendix A item 4 (2x2 box mean, u16 sum, truncating >> 2; level i is rows>>1, cols>>1 of
// level i-1; levels >= 1 packed consecutively in the pyramid buffer).
#ifndef _ORACLE_SHIM_DATATYPE_IMAGE_PYRAMID_H_
#define _ORACLE_SHIM_DATATYPE_IMAGE_PYRAMID_H_

#include "datatype_image.h"

class ImagePyramid {
public:
    static constexpr uint32_t kPyramidMaxLevel = 10;

    ImagePyramid() = default;
    ~ImagePyramid() {
        if (own_buff_ && buff_ != nullptr) {
            std::free(buff_);
        }
    }
    ImagePyramid(const ImagePyramid &) = delete;
    ImagePyramid &operator=(const ImagePyramid &) = delete;

    void SetPyramidBuff(uint8_t *buff, bool own) {
        buff_ = buff;
        own_buff_ = own;
    }
    void SetRawImage(uint8_t *data, int32_t rows, int32_t cols) { images_[0].SetImage(data, rows, cols); }

    bool CreateImagePyramid(uint32_t level) {
        if (images_[0].data() == nullptr || buff_ == nullptr) {
            return false;
        }
        level_ = level < kPyramidMaxLevel ? level : kPyramidMaxLevel;
        uint8_t *buf = buff_;
        for (uint32_t i = 1; i < level_; ++i) {
            const GrayImage &src = images_[i - 1];
            const int32_t rows = src.rows() >> 1;
            const int32_t cols = src.cols() >> 1;
            images_[i].SetImage(buf, rows, cols);
            buf += rows * cols;
            for (int32_t row = 0; row < rows; ++row) {
                for (int32_t col = 0; col < cols; ++col) {
                    const int32_t r2 = row << 1;
                    const int32_t c2 = col << 1;
                    const uint16_t sum = static_cast<uint16_t>(src.GetPixelValueNoCheck(r2, c2)) + static_cast<uint16_t>(src.GetPixelValueNoCheck(r2 + 1, c2)) +
                                         static_cast<uint16_t>(src.GetPixelValueNoCheck(r2, c2 + 1)) +
                                         static_cast<uint16_t>(src.GetPixelValueNoCheck(r2 + 1, c2 + 1));
                    images_[i].SetPixelValueNoCheck(row, col, static_cast<uint8_t>(sum >> 2));
                }
            }
        }
        return true;
    }

    // Shim-only extension used by the oracle driver: adopt externally built levels verbatim.
    void SetLevels(uint32_t level, uint8_t *const *data, const int32_t *rows, const int32_t *cols) {
        level_ = level;
        for (uint32_t i = 0; i < level; ++i) images_[i].SetImage(data[i], rows[i], cols[i]);
    }

    uint32_t level() const { return level_; }
    uint8_t *data() const { return images_[0].data(); }  // dense_optical_flow.cpp:38-39 only tests it against nullptr
    const GrayImage &GetImageConst(uint32_t i) const { return images_[i]; }
    GrayImage &GetImage(uint32_t i) { return images_[i]; }

private:
    GrayImage images_[kPyramidMaxLevel];
    uint8_t *buff_ = nullptr;
    bool own_buff_ = false;
    uint32_t level_ = 0;
};

#endif
