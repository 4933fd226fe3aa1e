// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/leaper.cuh"

namespace pg {
const GameVTable *pg_vtable_leaper() {
    static const GameVTable vt = make_vtable<LeaperGame>(GAME_LEAPER);
    return &vt;
}
}  // namespace pg
