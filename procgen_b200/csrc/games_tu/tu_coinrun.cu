// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/coinrun.cuh"

namespace pg {
const GameVTable *pg_vtable_coinrun() {
    static const GameVTable vt = make_vtable<CoinRun>(GAME_COINRUN);
    return &vt;
}
}  // namespace pg
