// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/plunder.cuh"

namespace pg {
const GameVTable *pg_vtable_plunder() {
    static const GameVTable vt = make_vtable<PlunderGame>(GAME_PLUNDER);
    return &vt;
}
}  // namespace pg
