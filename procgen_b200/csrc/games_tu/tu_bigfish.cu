// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/bigfish.cuh"

namespace pg {
const GameVTable *pg_vtable_bigfish() {
    static const GameVTable vt = make_vtable<BigFish>(GAME_BIGFISH);
    return &vt;
}
}  // namespace pg
