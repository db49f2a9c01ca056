// empty stand-in: the adapter does not need igl/get_seconds.h (test infrastructure, see Eigen/Core here)
