// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/dodgeball.cuh"

namespace pg {
const GameVTable *pg_vtable_dodgeball() {
    static const GameVTable vt = make_vtable<DodgeballGame>(GAME_DODGEBALL);
    return &vt;
}
}  // namespace pg
