/* vtkImageData with the calls of SobFuApp::save_field (src/apps/demo.cpp:252-284): a dense block of scalars */
#pragma once
#include <vector>
#ifndef VTK_FLOAT
#define VTK_FLOAT 10
#define VTK_DOUBLE 11
#endif
class vtkImageData {
public:
    void SetDimensions(int x, int y, int z) { dims_[0] = x; dims_[1] = y; dims_[2] = z; }
    int *GetDimensions() { return dims_; }
    void AllocateScalars(int type, int components) {
        type_ = type; comps_ = components;
        buf_.assign((size_t)dims_[0] * dims_[1] * dims_[2] * components * (type == VTK_DOUBLE ? 8 : 4), 0);
    }
    void *GetScalarPointer() { return buf_.data(); }
    int GetNumberOfScalarComponents() const { return comps_; }
    int GetScalarType() const { return type_; }
private:
    int dims_[3] = {0, 0, 0}, type_ = VTK_FLOAT, comps_ = 1;
    std::vector<unsigned char> buf_;
};
