/* vtkXMLImageDataWriter::{SetFileName, SetInputData, Write} (src/apps/demo.cpp:277-281): raw appended .vti, float scalars */
#pragma once
#include <sobfu_b200_io.hpp>
#include <vtkImageData.h>
#include <string>
class vtkXMLImageDataWriter {
public:
    void SetFileName(const char *name) { name_ = name ? name : ""; }
    void SetInputData(vtkImageData *im) { im_ = im; }
    int Write() {
        if (!im_ || name_.empty() || im_->GetScalarType() != VTK_FLOAT) return 0;
        const int *d = im_->GetDimensions();
        try {
            sobfu_b200::io::write_vti(name_, static_cast<const float *>(im_->GetScalarPointer()), d[0], d[1], d[2], im_->GetNumberOfScalarComponents());
        } catch (const std::exception &) {
            return 0;
        }
        return 1;
    }
private:
    std::string name_;
    vtkImageData *im_ = nullptr;
};
