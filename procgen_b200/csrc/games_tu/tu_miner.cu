// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/miner.cuh"

namespace pg {
const GameVTable *pg_vtable_miner() {
    static const GameVTable vt = make_vtable<MinerGame>(GAME_MINER);
    return &vt;
}
}  // namespace pg
