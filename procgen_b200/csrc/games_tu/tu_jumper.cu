// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/jumper.cuh"

namespace pg {
const GameVTable *pg_vtable_jumper() {
    static const GameVTable vt = make_vtable<JumperGame>(GAME_JUMPER);
    return &vt;
}
}  // namespace pg
