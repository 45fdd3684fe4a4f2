#include <torch/extension.h>
void register_pointwise(pybind11::module& m) {}
