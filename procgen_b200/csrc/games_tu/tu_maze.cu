// Translation unit of one game: instantiates its logic / render kernels and exports its dispatch entry.
#include "../pg_launch.cuh"
#include "../games/maze.cuh"

namespace pg {
const GameVTable *pg_vtable_maze() {
    static const GameVTable vt = make_vtable<MazeGame>(GAME_MAZE);
    return &vt;
}
}  // namespace pg
