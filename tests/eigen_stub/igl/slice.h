// empty stand-in: the adapter does not need igl/slice.h (test infrastructure, see Eigen/Core here)
