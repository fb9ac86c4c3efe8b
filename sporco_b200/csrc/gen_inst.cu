// gen_inst.cu -- explicit instantiation of the any-size launchers (direct-DFT path).
#include "launchers_impl.cuh"

namespace spcsc {

#define SPCSC_GEN_INST(T)                                                                           \
    template cudaError_t row_fwd_gen_launch<T>(const GenRowArgs<T>&, const T*, const T*,            \
                                               const AdmmState<T>*, C2<T>*);                        \
    template cudaError_t row_inv_gen_launch<T>(const GenRowArgs<T>&, const C2<T>*, T*, T);          \
    template cudaError_t row_inv_prox_gen_launch<T>(const GenRowArgs<T>&, const ProxArgs<T>&,       \
                                                    const C2<T>*, T*, T*, const AdmmState<T>*);     \
    template cudaError_t row_inv_prox_fwd_gen_launch<T>(const GenRowArgs<T>&, const PgmRowArgs<T>&, \
                                                        C2<T>*, T*);                                \
    template cudaError_t col_launch<T, 0>(int, ColLaunch<T>);

SPCSC_GEN_INST(float)
SPCSC_GEN_INST(double)

}  // namespace spcsc
