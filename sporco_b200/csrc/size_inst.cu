// size_inst.cu -- explicit instantiation of every size-templated launcher for one transform
// length.  Compiled once per supported length with -DSPCSC_SIZE=<n> (see build.py).
#include "launchers_impl.cuh"

#ifndef SPCSC_SIZE
#error "compile with -DSPCSC_SIZE=<power of two>"
#endif

namespace spcsc {

#define SPCSC_INST(T)                                                                          \
    template cudaError_t row_fwd_launch<T, SPCSC_SIZE>(const RowArgs<T>&, const T*, const T*,  \
                                                       const AdmmState<T>*, C2<T>*);           \
    template cudaError_t row_inv_launch<T, SPCSC_SIZE>(const RowArgs<T>&, const C2<T>*, T*, T); \
    template cudaError_t row_inv_prox_launch<T, SPCSC_SIZE>(const RowArgs<T>&,                 \
                                                            const ProxArgs<T>&, const C2<T>*,  \
                                                            T*, T*, const AdmmState<T>*);      \
    template cudaError_t row_inv_prox_fwd_launch<T, SPCSC_SIZE>(const RowArgs<T>&,             \
                                                                const PgmRowArgs<T>&, C2<T>*, T*); \
    template cudaError_t col_launch<T, SPCSC_SIZE>(int, ColLaunch<T>);

SPCSC_INST(float)
SPCSC_INST(double)

// kernel set v2 (float only)
template cudaError_t row_fwd2_launch<float, SPCSC_SIZE>(const RowArgs<float>&, const float*,
                                                        const float*, const AdmmState<float>*,
                                                        C2<float>*, const C2<float>*, int);
template cudaError_t row_inv_prox2_launch<float, SPCSC_SIZE>(const RowArgs<float>&,
                                                             const ProxArgs<float>&,
                                                             const C2<float>*, float*, float*,
                                                             const AdmmState<float>*,
                                                             const C2<float>*);
template cudaError_t col2_launch<float, SPCSC_SIZE>(int, ColLaunch<float>, const C2<float>*);

}  // namespace spcsc
