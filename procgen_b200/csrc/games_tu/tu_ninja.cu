// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/ninja.cuh"

namespace pg {
const GameVTable *pg_vtable_ninja() {
    static const GameVTable vt = make_vtable<Ninja>(GAME_NINJA);
    return &vt;
}
}  // namespace pg
