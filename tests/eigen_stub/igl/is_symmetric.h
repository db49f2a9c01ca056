// empty stand-in: the adapter does not need igl/is_symmetric.h (test infrastructure, see Eigen/Core here)
