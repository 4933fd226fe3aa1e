// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/bossfight.cuh"

namespace pg {
const GameVTable *pg_vtable_bossfight() {
    static const GameVTable vt = make_vtable<BossfightGame>(GAME_BOSSFIGHT);
    return &vt;
}
}  // namespace pg
