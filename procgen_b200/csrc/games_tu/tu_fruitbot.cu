// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/fruitbot.cuh"

namespace pg {
const GameVTable *pg_vtable_fruitbot() {
    static const GameVTable vt = make_vtable<FruitBotGame>(GAME_FRUITBOT);
    return &vt;
}
}  // namespace pg
