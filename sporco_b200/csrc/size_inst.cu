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

// kernel set v2 (register plans; float64 with half the elements per lane)
#define SPCSC_INST2(T)                                                                         \
    template cudaError_t row_fwd2_launch<T, SPCSC_SIZE>(const RowArgs<T>&, const T*, const T*, \
                                                        const AdmmState<T>*, C2<T>*,           \
                                                        const C2<T>*, int);                    \
    template cudaError_t row_inv_prox2_launch<T, SPCSC_SIZE>(const RowArgs<T>&,                \
                                                             const ProxArgs<T>&, const C2<T>*, \
                                                             T*, T*, const AdmmState<T>*,      \
                                                             const C2<T>*);                    \
    template cudaError_t col2_launch<T, SPCSC_SIZE>(int, ColLaunch<T>, const C2<T>*);

SPCSC_INST2(float)
SPCSC_INST2(double)

}  // namespace spcsc
