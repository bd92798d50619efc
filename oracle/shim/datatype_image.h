// ORACLE SHIM (test infrastructure).  Stand-in for Slam_Utility's `datatype_image.h` (absent from
// /root/reference); semantics frozen in SURVEY.md Appendix A item 3.
#ifndef _ORACLE_SHIM_DATATYPE_IMAGE_H_
#define _ORACLE_SHIM_DATATYPE_IMAGE_H_

#include "basic_type.h"

class GrayImage {
public:
    GrayImage() = default;
    GrayImage(uint8_t *data, int32_t rows, int32_t cols): data_(data), rows_(rows), cols_(cols) {}

    void SetImage(uint8_t *data, int32_t rows, int32_t cols) {
        data_ = data;
        rows_ = rows;
        cols_ = cols;
    }
    uint8_t *data() const { return data_; }
    int32_t rows() const { return rows_; }
    int32_t cols() const { return cols_; }

    // Integer access, no bounds test.
    inline uint8_t GetPixelValueNoCheck(int32_t row, int32_t col) const { return data_[row * cols_ + col]; }
    inline void SetPixelValueNoCheck(int32_t row, int32_t col, uint8_t v) { data_[row * cols_ + col] = v; }

    // Bilinear sample, no bounds test: base pixel by truncation, fractions by floor, four weighted terms
    // summed left to right in fp32.
    inline float GetPixelValueNoCheck(float row, float col) const {
        const uint8_t *values = &data_[static_cast<int32_t>(row) * cols_ + static_cast<int32_t>(col)];
        const float sub_row = row - std::floor(row);
        const float sub_col = col - std::floor(col);
        const float inv_sub_row = 1.0f - sub_row;
        const float inv_sub_col = 1.0f - sub_col;
        return inv_sub_col * inv_sub_row * static_cast<float>(values[0]) + sub_col * inv_sub_row * static_cast<float>(values[1]) +
               inv_sub_col * sub_row * static_cast<float>(values[cols_]) + sub_col * sub_row * static_cast<float>(values[cols_ + 1]);
    }

    // Bilinear sample with bounds test; false iff the position is outside [0, cols-1] x [0, rows-1].
    inline bool GetPixelValue(float row, float col, float *value) const {
        if (col < 0 || row < 0 || col > cols_ - 1 || row > rows_ - 1) {
            return false;
        }
        *value = GetPixelValueNoCheck(row, col);
        return true;
    }

private:
    uint8_t *data_ = nullptr;
    int32_t rows_ = 0;
    int32_t cols_ = 0;
};

#endif
