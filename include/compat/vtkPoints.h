/* included by the sobfu application (src/apps/demo.cpp:19-28), nothing of it is used in the headless build */
#pragma once
class vtkPoints {};
