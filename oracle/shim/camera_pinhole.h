// ORACLE SHIM (test infrastructure).  Stand-in for Sensor_Model's `camera_pinhole.h`, absent from /root/reference.
// direct_method_tracker.cpp:118,143 only constructs the camera and lifts a normalised-plane point to the image plane;
// a pinhole camera without distortion does u = fx * x + cx, v = fy * y + cy.
#ifndef _ORACLE_SHIM_CAMERA_PINHOLE_H_
#define _ORACLE_SHIM_CAMERA_PINHOLE_H_

#include "basic_type.h"

namespace sensor_model {

class CameraPinhole {
public:
    CameraPinhole(float fx, float fy, float cx, float cy, int32_t rows, int32_t cols) : fx_(fx), fy_(fy), cx_(cx), cy_(cy) {
        (void)rows;
        (void)cols;
    }
    void LiftFromNormalizedPlaneToImagePlane(const Vec2 &norm_xy, Vec2 &pixel_uv) const {
        pixel_uv.x() = fx_ * norm_xy.x() + cx_;
        pixel_uv.y() = fy_ * norm_xy.y() + cy_;
    }

private:
    float fx_, fy_, cx_, cy_;
};

}  // namespace sensor_model

#endif
