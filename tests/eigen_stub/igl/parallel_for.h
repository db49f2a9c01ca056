// empty stand-in: the adapter does not need igl/parallel_for.h (test infrastructure, see Eigen/Core here)
