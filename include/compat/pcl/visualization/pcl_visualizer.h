/* Headless stand-in for pcl::visualization::PCLVisualizer with the calls the sobfu application makes
 * (src/apps/demo.cpp:372-500).  sobfu_b200 renders nothing: every call is accepted and ignored, the first one says so once.
 * With PCL + VTK installed put them first on the include path to get the reference's viewer back. */
#pragma once
#include <pcl/PolygonMesh.h>

#include <cstdio>
#include <string>

namespace pcl {
namespace visualization {
class PCLVisualizer {
public:
    explicit PCLVisualizer(const std::string & = "", bool = true) { note(); }
    void createViewPort(double, double, double, double, int &viewport) { static int next = 1; viewport = next++; }
    bool addText(const std::string &, int, int, int, double, double, double, const std::string & = "", int = 0) { return true; }
    bool updateText(const std::string &, int, int, int, double, double, double, const std::string & = "") { return true; }
    void setCameraPosition(double, double, double, double, double, double, int = 0) {}
    bool addPolygonMesh(const pcl::PolygonMesh &, const std::string & = "polygon", int = 0) { return true; }
    bool updatePolygonMesh(const pcl::PolygonMesh &, const std::string & = "polygon") { return true; }
    void spinOnce(int = 1, bool = false) {}
    void saveScreenshot(const std::string &) {}
    bool wasStopped() const { return false; }
    void close() {}

private:
    static void note() {
        static bool said = false;
        if (!said) std::fprintf(stderr, "sobfu_b200: headless build -- the visualiser calls are ignored (meshes are still written with --enable-log)\n");
        said = true;
    }
};
}  // namespace visualization
}  // namespace pcl
