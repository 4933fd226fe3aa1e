// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/chaser.cuh"

namespace pg {
const GameVTable *pg_vtable_chaser() {
    static const GameVTable vt = make_vtable<ChaserGame>(GAME_CHASER);
    return &vt;
}
}  // namespace pg
