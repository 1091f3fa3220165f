// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Python module wrapper for the UNMODIFIED reference slam_ext sources.  The reference registers its ops
// through `pybind_slam_ext(py::module&)` (csrc/slam_ext/slam.cpp:31-37), which csrc/bind.cpp:38-39 attaches
// as the `slam_ext` submodule of one big module that also needs six unrelated extensions.  This file
// attaches only that one submodule so `oracle/_ref/vipe_ref_ext.so` exposes `slam_ext.ba`, `projmap`,
// `frame_distance`, `depth_filter`, `iproj` exactly as `vipe.ext.slam_ext` would.
#include <torch/extension.h>

void pybind_slam_ext(py::module &m);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    py::module m_slam = m.def_submodule("slam_ext");
    pybind_slam_ext(m_slam);
}
