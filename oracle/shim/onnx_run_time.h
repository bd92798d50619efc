// ORACLE SHIM (test infrastructure).  Stand-in for Slam_Utility's `onnx_run_time.h` + the ONNX Runtime C++ API, which are
// absent here.  Purpose: compile the reference's src/nn_feature_matcher/nn_feature_matcher.cpp WHERE IT LIES, unmodified, so that
// the post-processing of NNFeatureMatcher::Match (nn_feature_matcher.cpp:150-219: row / column arg-max, kMinValidMatchScore
// gate, mutual check; and the "matches" branch :160-178) runs as the reference wrote it.  There is no network here: the stub
// session returns whatever output tensors the test driver injected (shim::NextOutputs()), so the LightGlue model itself stays out
// of scope -- only what the reference does WITH its output is exercised.
#ifndef _ORACLE_SHIM_ONNX_RUN_TIME_H_
#define _ORACLE_SHIM_ONNX_RUN_TIME_H_

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "basic_type.h"

// ---- image-style dynamic matrices (Slam_Utility: TMatImg<T> = row-major Eigen matrix) -----------------------------------------
template <typename T>
struct TMatImg {
    std::vector<T> v;
    int r = 0, c = 0;
    struct RowRef {
        TMatImg *m;
        int i;
        template <int N>
        RowRef &operator=(const shim::Mat<1, N> &row) {
            for (int j = 0; j < N && j < m->c; ++j) m->v[static_cast<size_t>(i) * m->c + j] = static_cast<T>(row(0, j));
            return *this;
        }
    };
    void setZero(size_t rows, size_t cols) {
        r = static_cast<int>(rows), c = static_cast<int>(cols);
        v.assign(rows * cols, T(0));
    }
    int rows() const { return r; }
    int cols() const { return c; }
    RowRef row(int i) { return RowRef{this, i}; }
    T &operator()(int i, int j) { return v[static_cast<size_t>(i) * c + j]; }
    const T &operator()(int i, int j) const { return v[static_cast<size_t>(i) * c + j]; }
};
using MatImgF = TMatImg<float>;

namespace Eigen {
template <typename M>
class Map;
// Read-only view of a row-major matrix: what nn_feature_matcher.cpp uses of Eigen::Map<const TMatImg<T>> (rows, cols, (i, j),
// col(j)(i), row(i)(j)).
template <typename T>
class Map<const TMatImg<T>> {
public:
    struct Slice {
        const T *p;
        int n, stride;
        uint32_t rows() const { return static_cast<uint32_t>(n); }
        uint32_t cols() const { return static_cast<uint32_t>(n); }
        const T &operator()(uint32_t k) const { return p[static_cast<size_t>(k) * stride]; }
    };
    Map(const T *data, int rows, int cols) : p_(data), r_(rows), c_(cols) {}
    uint32_t rows() const { return static_cast<uint32_t>(r_); }
    uint32_t cols() const { return static_cast<uint32_t>(c_); }
    const T &operator()(int i, int j) const { return p_[static_cast<size_t>(i) * c_ + j]; }
    Slice col(uint32_t j) const { return Slice{p_ + j, r_, c_}; }
    Slice row(uint32_t i) const { return Slice{p_ + static_cast<size_t>(i) * c_, c_, 1}; }

private:
    const T *p_;
    int r_, c_;
};
}  // namespace Eigen

// Feature_Detector's NN descriptor types: fixed-size float column vectors (256-d per BASELINE.json); only default construction,
// rows() and transpose() are used by nn_feature_matcher.cpp.
using SuperpointDescriptorType = shim::Mat<256, 1>;
using DiskDescriptorType = shim::Mat<128, 1>;

// ---- the slice of the ONNX Runtime C++ API the file touches ---------------------------------------------------------------
enum OrtLoggingLevel { ORT_LOGGING_LEVEL_WARNING = 2 };
enum GraphOptimizationLevel { ORT_ENABLE_EXTENDED = 2 };
enum ExecutionMode { ORT_PARALLEL = 1 };
enum OrtAllocatorType { OrtDeviceAllocator = 0 };
enum OrtMemType { OrtMemTypeDefault = 0 };

namespace Ort {
struct Exception : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct Env {
    Env(OrtLoggingLevel, const char *) {}
};
struct SessionOptions {
    void SetGraphOptimizationLevel(GraphOptimizationLevel) {}
    void SetExecutionMode(ExecutionMode) {}
};
struct MemoryInfo {
    MemoryInfo(std::nullptr_t) {}
    MemoryInfo() = default;
    static MemoryInfo CreateCpu(OrtAllocatorType, OrtMemType) { return MemoryInfo(); }
};
struct RunOptions {
    void SetRunLogVerbosityLevel(int) {}
};
// A tensor: either int64 or float payload, rows x cols.
struct Value {
    Value(std::nullptr_t) {}
    Value() = default;
    std::vector<float> f;
    std::vector<int64_t> i64;
    int rows = 0, cols = 0;
};
}  // namespace Ort

namespace shim {
// The output tensors the next Session::Run returns (set by the test driver).
inline std::vector<Ort::Value> &NextOutputs() {
    static std::vector<Ort::Value> outputs;
    return outputs;
}
}  // namespace shim

namespace Ort {
struct Session {
    Session(std::nullptr_t) {}
    Session(Env &, const char *, const SessionOptions &) : live_(true) {}
    explicit operator bool() const { return live_; }
    bool operator!() const { return !live_; }
    std::vector<Value> Run(const RunOptions &, const char *const *, const Value *, size_t, const char *const *, size_t) { return shim::NextOutputs(); }

private:
    bool live_ = false;
};
}  // namespace Ort

// Slam_Utility's helper class around the runtime.
class OnnxRuntime {
public:
    static void TryToEnableCuda(Ort::SessionOptions &) {}
    static void ReportInformationOfSession(const Ort::Session &) {}
    static void ReportInformationOfOrtValue(const Ort::Value &) {}
    // LightGlue takes four inputs (keypoints x2, descriptors x2); the score-matrix models have one output, the fused ones two.
    static void GetSessionIO(const Ort::Session &, std::vector<std::string> &inputs, std::vector<std::string> &outputs) {
        inputs = {"kpts0", "kpts1", "desc0", "desc1"};
        outputs = {"scores"};
    }
    static void ConvertMatrixToTensor(const MatImgF &, const Ort::MemoryInfo &, Ort::Value &) {}
    static void ConvertTensorToImageMatrice(const Ort::Value &t, std::vector<Eigen::Map<const TMatImg<float>>> &out) {
        out.clear();
        out.emplace_back(t.f.data(), t.rows, t.cols);
    }
    static void ConvertTensorToImageMatrice(const Ort::Value &t, std::vector<Eigen::Map<const TMatImg<int64_t>>> &out) {
        out.clear();
        out.emplace_back(t.i64.data(), t.rows, t.cols);
    }
};

#endif
