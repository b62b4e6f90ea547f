// sobfu_headless -- the sobfu application without a display: reads a sequence of 16-bit depth
// PNGs (optionally masked by <dir>/omask) and the reference's .ini parameter files, runs SobFusion::operator() per frame and
// writes the canonical / warped meshes as legacy VTK files.  Command line, directory layout, parameter handling and console
// output follow src/apps/demo.cpp of dgrzech/sobfu (SobFuApp, lines 30-519; parse_flags / main, lines 526-618); what it adds
// is a synthetic depth source, a frame limit and a machine-readable summary, what it drops is the PCL viewer.
//
//   sobfu_headless [OPTIONS] <file path> <ini path>
//     --enable-log            save canonical_mesh_XXXXXX.vtk / canonical_warped_to_live_mesh_XXXXXX.vtk into <file path>/meshes
//     --verbose / --vverbose  solver verbosity 1 / 2
//     --enable-viz, --enable-viz-detailed   accepted; there is no viewer in this build (meshes are extracted as the viewer would)
//     --synthetic N           no input directory: N frames of an analytically ray-cast sphere (640x480, moving 2 mm / frame)
//     --frames K              stop after K frames
//     --out DIR               mesh directory (default <file path>/meshes; required with --synthetic and --enable-log)
//     --save-field            with --enable-log: also field_XXXXXX.vti (the deformation field, demo.cpp:252-284)
//     --device D              CUDA device (default 0)
//     --json                  one JSON line at the end: frames, seconds, frames/s, vertices of the last canonical mesh
// The host side is C++ over the C ABI (include/sobfu_b200.h) through the drop-in headers (include/sobfu/*.hpp).
#include <sobfu/sob_fusion.hpp>

#include <boost/program_options.hpp>
#include <opencv2/highgui/highgui.hpp>
#include <pcl/io/vtk_io.h>

#include <chrono>
#include <cmath>
#include <iostream>

namespace po = boost::program_options;

struct Options {
    std::string file_path, params_path, out_path;
    bool logger = false, viz = false, verbose = false, vverbose = false, save_field = false, json = false;
    int synthetic = 0, max_frames = -1, device = 0;
};

static void usage() {
    std::cout << "USAGE: sobfu_headless [OPTIONS] <file path> <ini path>\n"
                 "\t--help -h:    display help\n\t--enable-log: log output meshes\n\t--verbose: low verbosity\n\t--vverbose: high verbosity\n"
                 "\t--enable-viz, --enable-viz-detailed: accepted, no viewer in this build\n"
                 "\t--synthetic N: N synthetic frames instead of <file path>/depth (then only <ini path> is positional)\n"
                 "\t--frames K: stop after K frames\n\t--out DIR: mesh directory\n\t--save-field: also save the deformation field (.vti)\n"
                 "\t--device D: CUDA device\n\t--json: print a JSON summary line\n";
}

// demo.cpp:84-160: the option set of the .ini files; TSDF_TRUNC_DIST / ETA / VOL_POSE_T_Z are read from the map afterwards
static void declare_parameters(po::options_description &desc, Params &params) {
    desc.add_options()("VOL_DIMS_X", po::value<int>(&params.volume_dims[0]), "no. of voxels along x axis");
    desc.add_options()("VOL_DIMS_Y", po::value<int>(&params.volume_dims[1]), "no. of voxels along y axis");
    desc.add_options()("VOL_DIMS_Z", po::value<int>(&params.volume_dims[2]), "no. of voxels along z axis");
    desc.add_options()("VOL_SIZE_X", po::value<float>(&params.volume_size[0]), "vol. size along x axis (metres)");
    desc.add_options()("VOL_SIZE_Y", po::value<float>(&params.volume_size[1]), "vol. size along y axis (metres)");
    desc.add_options()("VOL_SIZE_Z", po::value<float>(&params.volume_size[2]), "vol. size along z axis (metres)");
    desc.add_options()("TSDF_TRUNC_DIST", po::value<float>(), "truncation distance (voxels)");
    desc.add_options()("ETA", po::value<float>(), "expected object thickness (voxels)");
    desc.add_options()("TSDF_MAX_WEIGHT", po::value<float>(&params.tsdf_max_weight), "max. tsdf weight");
    desc.add_options()("GRADIENT_DELTA_FACTOR", po::value<float>(&params.gradient_delta_factor), "delta factor of the tsdf gradient (voxels)");
    desc.add_options()("INTR_FX", po::value<float>(&params.intr.fx), "focal length x");
    desc.add_options()("INTR_FY", po::value<float>(&params.intr.fy), "focal length y");
    desc.add_options()("INTR_CX", po::value<float>(&params.intr.cx), "principal point x");
    desc.add_options()("INTR_CY", po::value<float>(&params.intr.cy), "principal point y");
    desc.add_options()("TRUNC_DEPTH", po::value<float>(&params.icp_truncate_depth_dist), "depth map truncation distance (metres)");
    desc.add_options()("VOL_POSE_T_Z", po::value<float>(), "camera to volume translation along z axis");
    desc.add_options()("BILATERAL_SIGMA_DEPTH", po::value<float>(&params.bilateral_sigma_depth), "bilateral filter sigma z");
    desc.add_options()("BILATERAL_SIGMA_SPATIAL", po::value<float>(&params.bilateral_sigma_spatial), "bilateral filter sigma x-y");
    desc.add_options()("BILATERAL_KERNEL_SIZE", po::value<int>(&params.bilateral_kernel_size), "bilateral filter kernel size");
    desc.add_options()("START_FRAME", po::value<int>(&params.start_frame), "frame when to start registration");
    desc.add_options()("MAX_ITER", po::value<int>(&params.max_iter), "max. no. of iterations of the solver");
    desc.add_options()("MAX_UPDATE_NORM", po::value<float>(&params.max_update_norm), "max. update norm when running the solver");
    desc.add_options()("S", po::value<int>(&params.s), "Sobolev kernel size");
    desc.add_options()("LAMBDA", po::value<float>(&params.lambda), "Sobolev filter parameter");
    desc.add_options()("ALPHA", po::value<float>(&params.alpha), "gradient descent step size");
    desc.add_options()("W_REG", po::value<float>(&params.w_reg), "regularisation weight");
}

static bool read_params(const std::string &path, Params &params) {
    po::options_description desc("parameters");
    declare_parameters(desc, params);
    po::variables_map vm;
    std::ifstream settings_file(path);
    if (!settings_file) {
        std::cerr << "error: cannot open '" << path << "'. exiting..." << std::endl;
        return false;
    }
    try {
        po::store(po::parse_config_file(settings_file, desc), vm);
        po::notify(vm);
        for (const char *need : {"VOL_DIMS_X", "VOL_DIMS_Y", "VOL_DIMS_Z", "VOL_SIZE_X", "VOL_SIZE_Y", "VOL_SIZE_Z", "TSDF_TRUNC_DIST", "ETA", "VOL_POSE_T_Z",
                                 "MAX_ITER", "S", "LAMBDA", "ALPHA", "W_REG"})
            if (!vm.count(need)) throw po::error(std::string("missing option '") + need + "'");
        // parameters stored in units of voxels (demo.cpp:68-74)
        params.tsdf_trunc_dist = vm["TSDF_TRUNC_DIST"].as<float>() * params.voxel_sizes()[0];
        params.eta = vm["ETA"].as<float>() * params.voxel_sizes()[0];
        params.volume_pose = cv::Affine3f().translate(cv::Vec3f(-params.volume_size[0] / 2.f, -params.volume_size[1] / 2.f, vm["VOL_POSE_T_Z"].as<float>()));
    } catch (const std::exception &e) {
        std::cerr << "error: " << path << ": " << e.what() << ". exiting..." << std::endl;
        return false;
    }
    return true;
}

// analytically ray-cast sphere of radius 0.15 m centred at (0.002 * frame, 0, 0.5) m seen through the .ini's intrinsics, ushort mm
static cv::Mat synthetic_depth(int frame, const Params &p) {
    cv::Mat depth(p.rows, p.cols, CV_16UC1);
    const double c[3] = {0.002 * frame, 0.0, 0.5}, radius = 0.15;
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2] - radius * radius;
    for (int v = 0; v < p.rows; ++v)
        for (int u = 0; u < p.cols; ++u) {
            const double dx = (u - p.intr.cx) / p.intr.fx, dy = (v - p.intr.cy) / p.intr.fy;
            const double a = dx * dx + dy * dy + 1.0, b = -2.0 * (dx * c[0] + dy * c[1] + c[2]), disc = b * b - 4 * a * cc;
            depth.ptr<unsigned short>(v)[u] = disc > 0 ? (unsigned short)std::lround((-b - std::sqrt(disc)) / (2 * a) * 1000.0) : 0;
        }
    return depth;
}

static std::string frame_name(int i) {
    std::stringstream ss;
    ss << std::setw(6) << std::setfill('0') << i;
    return ss.str();
}
static int no_vertices(const pcl::PolygonMesh::Ptr &mesh) { return mesh->cloud.point_step ? (int)(mesh->cloud.data.size() / mesh->cloud.point_step) : 0; }

int main(int argc, char *argv[]) {
    Options o;
    std::vector<std::string> positional;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next_int = [&](int &dst) { if (i + 1 < argc) dst = std::atoi(argv[++i]); };
        if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (a == "--enable-log") o.logger = true;
        else if (a == "--enable-viz" || a == "--enable-viz-detailed") o.viz = true;
        else if (a == "--verbose") o.verbose = true;
        else if (a == "--vverbose") o.vverbose = true;
        else if (a == "--save-field") o.save_field = true;
        else if (a == "--json") o.json = true;
        else if (a == "--synthetic") next_int(o.synthetic);
        else if (a == "--frames") next_int(o.max_frames);
        else if (a == "--device") next_int(o.device);
        else if (a == "--out") { if (i + 1 < argc) o.out_path = argv[++i]; }
        else positional.push_back(a);
    }
    if (o.synthetic > 0 && positional.size() == 1) o.params_path = positional[0];
    else if (positional.size() >= 2) { o.file_path = positional[0]; o.params_path = positional[1]; }
    else {
        std::cerr << "error: incorrect number of arguments; please supply path to source data and .ini file; exiting..." << std::endl;
        return -1;
    }

    kfusion::cuda::setDevice(o.device);
    kfusion::cuda::printShortCudaDeviceInfo(o.device);

    Params params;
    params.verbosity = o.verbose ? 1 : (o.vverbose ? 2 : 0);      // demo.cpp:47-51
    if (!read_params(o.params_path, params)) return 1;

    // input (demo.cpp:176-199): <file path>/depth is required, <file path>/omask optional; colour frames are not needed here
    std::vector<cv::String> depths, masks;
    if (o.synthetic <= 0) {
        if (!boost::filesystem::exists(o.file_path)) {
            std::cerr << "error: directory '" << o.file_path << "' does not exist. exiting" << std::endl;
            return 1;
        }
        if (!boost::filesystem::exists(o.file_path + "/depth")) {
            std::cerr << "error: source directory should contain a 'depth' folder. exiting..." << std::endl;
            return 1;
        }
        cv::glob(o.file_path + "/depth", depths);
        std::sort(depths.begin(), depths.end());
        if (boost::filesystem::exists(o.file_path + "/omask")) {
            cv::glob(o.file_path + "/omask", masks);
            std::sort(masks.begin(), masks.end());
        }
    }
    const bool has_masks = !masks.empty();
    size_t n_frames = o.synthetic > 0 ? (size_t)o.synthetic : depths.size();
    if (o.max_frames >= 0 && (size_t)o.max_frames < n_frames) n_frames = (size_t)o.max_frames;

    if (o.out_path.empty()) o.out_path = o.file_path.empty() ? std::string("meshes") : o.file_path + "/meshes";
    if (o.logger && boost::filesystem::create_directory(o.out_path)) std::cout << "created output directory for meshes" << std::endl;

    SobFusion sobfu(params);
    kfusion::cuda::Depth depth_device;
    double time_ms = 0.0, total_ms = 0.0;
    int last_vertices = 0;
    for (size_t i = 0; i < n_frames; ++i) {
        cv::Mat depth = o.synthetic > 0 ? synthetic_depth((int)i, params) : cv::imread(depths[i], CV_LOAD_IMAGE_ANYDEPTH);
        if (!depth.data || depth.type() != CV_16UC1) {
            std::cerr << "error: image could not be read; check for improper permissions or invalid formats. exiting..." << std::endl;
            return 1;
        }
        if (has_masks && i < masks.size()) {                         // demo.cpp:304-308
            cv::Mat mask = cv::imread(masks[i], CV_8U), depth_masked = cv::Mat::zeros(depth.size(), depth.type());
            if (!mask.data || !(mask.size() == depth.size())) {
                std::cerr << "error: mask could not be read or does not match the depth map. exiting..." << std::endl;
                return 1;
            }
            depth.copyTo(depth_masked, mask);
            depth = depth_masked;
        }
        depth_device.upload(depth.data, depth.step, depth.rows, depth.cols);
        const auto t0 = std::chrono::steady_clock::now();
        {
            kfusion::SampledScopeTime fps(time_ms);
            sobfu(depth_device);
        }
        total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

        if (o.logger || o.viz) {                                     // demo.cpp:340-371
            pcl::PolygonMesh::Ptr mesh_global = sobfu.get_phi_global_mesh();
            last_vertices = no_vertices(mesh_global);
            std::cout << "no. of point-normal pairs in the canonical model: " << last_vertices << std::endl;
            pcl::PolygonMesh::Ptr mesh_global_psi_inv;
            if (i >= 1) {
                mesh_global_psi_inv = sobfu.get_phi_global_psi_inv_mesh();
                std::cout << "no. of point-normal pairs in the canonical model warped to live: " << no_vertices(mesh_global_psi_inv) << std::endl;
            }
            if (o.logger) {
                const std::string num = frame_name((int)i);
                if (pcl::io::saveVTKFile(o.out_path + "/canonical_mesh_" + num + ".vtk", *mesh_global) == 0) std::cout << "saved canonical_mesh_" + num + ".vtk" << std::endl;
                if (i >= 1 && pcl::io::saveVTKFile(o.out_path + "/canonical_warped_to_live_mesh_" + num + ".vtk", *mesh_global_psi_inv) == 0)
                    std::cout << "saved canonical_warped_to_live_mesh_" + num + ".vtk" << std::endl;
                if (o.save_field) {
                    std::shared_ptr<sobfu::cuda::DeformationField> psi = sobfu.getDeformationField();
                    std::vector<float> host((size_t)params.volume_dims[0] * params.volume_dims[1] * params.volume_dims[2] * 4);
                    kfusion::cuda::CudaData data = psi->get_data();
                    data.download(host.data());
                    sobfu_b200::io::write_vti(o.out_path + "/field_" + num + ".vti", host.data(), params.volume_dims[0], params.volume_dims[1], params.volume_dims[2], 4);
                    std::cout << "saved the vector field to .vti" << std::endl;
                }
            }
        }
    }
    if (o.json)
        std::printf("{\"frames\": %zu, \"seconds\": %.6f, \"frames_per_s\": %.4f, \"vertices\": %d, \"volume\": [%d, %d, %d]}\n", n_frames, total_ms / 1e3,
                    total_ms > 0 ? n_frames / (total_ms / 1e3) : 0.0, last_vertices, params.volume_dims[0], params.volume_dims[1], params.volume_dims[2]);
    return 0;
}
