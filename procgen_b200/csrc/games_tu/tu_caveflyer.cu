// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/caveflyer.cuh"

namespace pg {
const GameVTable *pg_vtable_caveflyer() {
    static const GameVTable vt = make_vtable<CaveFlyerGame>(GAME_CAVEFLYER);
    return &vt;
}
}  // namespace pg
