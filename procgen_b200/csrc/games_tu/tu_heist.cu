// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/heist.cuh"

namespace pg {
const GameVTable *pg_vtable_heist() {
    static const GameVTable vt = make_vtable<HeistGame>(GAME_HEIST);
    return &vt;
}
}  // namespace pg
