/* vtkSmartPointer<T>::New() / operator-> for the two VTK classes the sobfu application touches (demo.cpp:252-284) */
#pragma once
#include <memory>
template <class T>
class vtkSmartPointer {
public:
    vtkSmartPointer() {}
    static vtkSmartPointer New() { vtkSmartPointer p; p.p_ = std::make_shared<T>(); return p; }
    T *operator->() const { return p_.get(); }
    T *Get() const { return p_.get(); }
    T *GetPointer() const { return p_.get(); }
    operator T *() const { return p_.get(); }
private:
    std::shared_ptr<T> p_;
};
