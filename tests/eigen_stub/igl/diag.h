// empty stand-in: the adapter does not need igl/diag.h (test infrastructure, see Eigen/Core here)
