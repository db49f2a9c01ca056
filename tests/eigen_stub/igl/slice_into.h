// empty stand-in: the adapter does not need igl/slice_into.h (test infrastructure, see Eigen/Core here)
