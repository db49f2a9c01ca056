// igl/get_seconds.h -- included by the reference headers, nothing of it is used on the hot path (test infrastructure shim)
