// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/climber.cuh"

namespace pg {
const GameVTable *pg_vtable_climber() {
    static const GameVTable vt = make_vtable<Climber>(GAME_CLIMBER);
    return &vt;
}
}  // namespace pg
