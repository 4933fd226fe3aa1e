// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/starpilot.cuh"

namespace pg {
const GameVTable *pg_vtable_starpilot() {
    static const GameVTable vt = make_vtable<StarpilotGame>(GAME_STARPILOT);
    return &vt;
}
}  // namespace pg
