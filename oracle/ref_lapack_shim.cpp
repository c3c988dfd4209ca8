// oracle/ref_lapack_shim.cpp -- TEST INFRASTRUCTURE: compiles the reference's OWN ?potrf_ / ?getrf_ (lapack/cholesky.cpp,
// lapack/lu.cpp) from where they lie under /root/reference, one scalar type per compilation (-DLAPACK_T=0..3), without the
// eigenvalue / SVD parts of lapack/{single,double,complex_single,complex_double}.cpp that would triple the build time.
// The defines mirror those four files (lapack/double.cpp:10-13, lapack/complex_double.cpp:10-14).
#include <complex>
#if LAPACK_T == 0
#define SCALAR float
#define SCALAR_SUFFIX s
#define SCALAR_SUFFIX_UP "S"
#define ISCOMPLEX 0
#elif LAPACK_T == 1
#define SCALAR double
#define SCALAR_SUFFIX d
#define SCALAR_SUFFIX_UP "D"
#define ISCOMPLEX 0
#elif LAPACK_T == 2
#define SCALAR std::complex<float>
#define SCALAR_SUFFIX c
#define SCALAR_SUFFIX_UP "C"
#define REAL_SCALAR_SUFFIX s
#define ISCOMPLEX 1
#else
#define SCALAR std::complex<double>
#define SCALAR_SUFFIX z
#define SCALAR_SUFFIX_UP "Z"
#define REAL_SCALAR_SUFFIX d
#define ISCOMPLEX 1
#endif
#include "cholesky.cpp"
#include "lu.cpp"
