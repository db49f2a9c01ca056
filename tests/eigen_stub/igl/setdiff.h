// empty stand-in: the adapter does not need igl/setdiff.h (test infrastructure, see Eigen/Core here)
