// Which image files each game binds to each object type, and which background group it uses.
// Data restated from every games/<name>.cpp `asset_for_type` / `load_background_images`
// (file:line cited per game). Host only.
#pragma once
#include "pg_assets_host.h"

namespace pg {
namespace host {

inline GameAssetNames game_asset_names(int game_id) {
    GameAssetNames g;
    auto &T = g.by_type;
    switch (game_id) {
    case GAME_COINRUN: {  // coinrun.cpp:59-121
        g.bg_group = "platform_backgrounds";
        const char *colors[5] = {"Beige", "Blue", "Green", "Pink", "Yellow"};
        const char *poses[4] = {"stand", "jump", "walk1", "walk2"};
        const int pose_type[4] = {0 /*PLAYER*/, 9 /*PLAYER_JUMP*/, 12 /*PLAYER_RIGHT1*/, 13 /*PLAYER_RIGHT2*/};
        for (int p = 0; p < 4; p++)
            for (auto col : colors)
                T[pose_type[p]].push_back(std::string("kenney/Players/128x256/") + col + "/alien" + col + "_" + poses[p] + ".png");
        const char *enemies[9] = {"slimeBlock", "slimePurple", "slimeBlue", "slimeGreen", "mouse", "snail", "ladybug", "wormGreen", "wormPink"};
        for (auto e : enemies) {
            T[6].push_back(std::string("kenney/Enemies/") + e + ".png");       // ENEMY1
            T[7].push_back(std::string("kenney/Enemies/") + e + "_move.png");  // ENEMY2
        }
        T[1] = {"kenney/Items/coinGold.png"};  // GOAL
        const char *grounds[6] = {"Dirt", "Grass", "Planet", "Sand", "Snow", "Stone"};
        for (auto gr : grounds) {
            T[16].push_back(std::string("kenney/Ground/") + gr + "/" + lower(gr) + "Mid.png");     // WALL_TOP
            T[15].push_back(std::string("kenney/Ground/") + gr + "/" + lower(gr) + "Center.png");  // WALL_MID
        }
        T[18] = {"kenney/Tiles/lavaTop_low.png"};  // LAVA_TOP
        T[17] = {"kenney/Tiles/lava.png"};         // LAVA_MID
        T[2] = {"kenney/Enemies/sawHalf.png"};
        T[3] = {"kenney/Enemies/sawHalf_move.png"};
        T[20] = {"kenney/Tiles/boxCrate.png", "kenney/Tiles/boxCrate_double.png", "kenney/Tiles/boxCrate_single.png",
                 "kenney/Tiles/boxCrate_warning.png"};
        break;
    }
    case GAME_BIGFISH:  // bigfish.cpp:31-43
        g.bg_group = "water_backgrounds";
        T[0] = {"misc_assets/fishTile_072.png"};
        T[2] = {"misc_assets/fishTile_074.png", "misc_assets/fishTile_078.png", "misc_assets/fishTile_080.png"};
        break;
    case GAME_FRUITBOT: {  // fruitbot.cpp:47-81
        g.bg_group = "topdown_backgrounds";
        T[0] = {"misc_assets/robot_3Dblue.png"};
        T[1] = {"misc_assets/tileStone_slope.png"};
        T[2] = {"misc_assets/tileStone_slope.png"};
        T[3] = {"misc_assets/keyRed2.png"};
        for (int i = 1; i <= 6; i++) {
            T[4].push_back("misc_assets/food" + std::to_string(i) + ".png");
            T[7].push_back("misc_assets/fruit" + std::to_string(i) + ".png");
        }
        T[10] = {"misc_assets/fenceYellow.png"};
        T[11] = {"misc_assets/lockRed2.png"};
        T[12] = {"misc_assets/present1.png", "misc_assets/present2.png", "misc_assets/present3.png"};
        break;
    }
    case GAME_HEIST:  // heist.cpp:36-60
        g.bg_group = "topdown_backgrounds";
        T[51] = {"kenney/Ground/Dirt/dirtCenter.png"};
        T[9] = {"misc_assets/gemYellow.png"};
        T[0] = {"misc_assets/spaceAstronauts_008.png"};
        T[2] = {"misc_assets/keyBlue.png", "misc_assets/keyGreen.png", "misc_assets/keyRed.png"};
        T[1] = {"misc_assets/lock_blue.png", "misc_assets/lock_green.png", "misc_assets/lock_red.png"};
        break;
    case GAME_MAZE:  // maze.cpp:26-38
        g.bg_group = "topdown_backgrounds";
        T[51] = {"kenney/Ground/Sand/sandCenter.png"};
        T[2] = {"misc_assets/cheese.png"};
        T[0] = {"kenney/Enemies/mouse_move.png"};
        break;
    case GAME_BOSSFIGHT:  // bossfight.cpp:77-109
        g.bg_group = "space_backgrounds";
        T[0] = {"misc_assets/playerShip1_blue.png", "misc_assets/playerShip1_green.png", "misc_assets/playerShip2_orange.png",
                "misc_assets/playerShip3_red.png"};
        T[2] = {"misc_assets/enemyShipBlack1.png", "misc_assets/enemyShipBlue2.png", "misc_assets/enemyShipGreen3.png",
                "misc_assets/enemyShipRed4.png"};
        T[4] = {"misc_assets/laserGreen14.png", "misc_assets/laserRed11.png", "misc_assets/laserBlue09.png"};
        T[1] = {"misc_assets/laserGreen14.png", "misc_assets/laserRed11.png", "misc_assets/laserBlue09.png"};
        T[3] = {"misc_assets/shield2.png"};
        T[7] = {"misc_assets/spaceMeteors_001.png", "misc_assets/spaceMeteors_002.png", "misc_assets/spaceMeteors_003.png",
                "misc_assets/spaceMeteors_004.png", "misc_assets/meteorGrey_big1.png", "misc_assets/meteorGrey_big2.png",
                "misc_assets/meteorGrey_big3.png", "misc_assets/meteorGrey_big4.png"};
        break;
    case GAME_CAVEFLYER:  // caveflyer.cpp:36-54
        g.bg_group = "space_backgrounds";
        T[1] = {"misc_assets/ufoGreen2.png"};
        T[2] = {"misc_assets/meteorBrown_big1.png"};
        T[3] = {"misc_assets/ufoRed2.png"};
        T[4] = {"misc_assets/laserBlue02.png"};
        T[5] = {"misc_assets/enemyShipBlue4.png"};
        T[0] = {"misc_assets/playerShip1_red.png"};
        T[8] = {"misc_assets/groundA.png"};
        T[9] = {"misc_assets/towerDefense_tile295.png"};
        break;
    case GAME_CHASER:  // chaser.cpp:51-75
        g.bg_group = "topdown_simple_backgrounds";
        T[0] = {"misc_assets/enemyFloating_1b.png"};
        T[6] = {"misc_assets/enemyFlying_1.png"};
        T[7] = {"misc_assets/enemyFlying_2.png"};
        T[8] = {"misc_assets/enemyFlying_3.png"};
        T[2] = {"misc_assets/yellowCrystal.png"};
        T[3] = {"misc_assets/enemyWalking_1b.png"};
        T[4] = {"misc_assets/enemySpikey_1b.png"};
        T[5] = {"misc_assets/tileStone_slope.png"};
        break;
    case GAME_CLIMBER: {  // climber.cpp:47-89
        g.bg_group = "platform_backgrounds";
        const char *cols[4] = {"Blue", "Green", "Grey", "Red"};
        for (auto col : cols) {
            T[0].push_back(std::string("platformer/player") + col + "_stand.png");
            T[9].push_back(std::string("platformer/player") + col + "_walk4.png");
            T[12].push_back(std::string("platformer/player") + col + "_walk1.png");
            T[13].push_back(std::string("platformer/player") + col + "_walk2.png");
        }
        T[16] = {"platformer/tileBlue_05.png", "platformer/tileGreen_05.png", "platformer/tileYellow_06.png", "platformer/tileBrown_06.png"};
        T[15] = {"platformer/tileBlue_08.png", "platformer/tileGreen_08.png", "platformer/tileYellow_09.png", "platformer/tileBrown_09.png"};
        T[6] = {"platformer/enemySwimming_1.png"};
        T[7] = {"platformer/enemySwimming_2.png"};
        T[1] = {"platformer/yellowCrystal.png"};
        break;
    }
    case GAME_LEAPER:  // leaper.cpp:42-69
        g.bg_group = "topdown_backgrounds";
        T[2] = {"misc_assets/roadTile6b.png"};
        T[3] = {"misc_assets/terrainTile6.png"};
        T[4] = {"misc_assets/car_yellow_5.png", "misc_assets/car_black_1.png", "misc_assets/car_blue_2.png",
                "misc_assets/car_green_3.png", "misc_assets/car_red_4.png"};
        T[1] = {"misc_assets/elementWood044.png"};
        T[0] = {"misc_assets/frog1.png", "misc_assets/frog2.png", "misc_assets/frog4.png", "misc_assets/frog6.png",
                "misc_assets/frog7.png"};
        T[5] = {"misc_assets/finish2.png"};
        break;
    case GAME_NINJA:  // ninja.cpp:46-75
        g.bg_group = "platform_backgrounds";
        T[20] = {"misc_assets/tile_bricksGrey.png", "misc_assets/tile_bricksGrown.png", "misc_assets/tile_bricksRed.png"};
        T[1] = {"platformer/shroom1.png", "platformer/shroom2.png", "platformer/shroom3.png", "platformer/shroom4.png",
                "platformer/shroom5.png", "platformer/shroom6.png"};
        T[0] = {"platformer/zombie_idle.png"};
        T[9] = {"platformer/zombie_jump.png"};
        T[12] = {"platformer/zombie_walk1.png"};
        T[13] = {"platformer/zombie_walk2.png"};
        T[6] = {"misc_assets/bomb.png"};
        T[7] = {"misc_assets/saw.png"};
        T[14] = {"misc_assets/bomb.png"};
        break;
    case GAME_PLUNDER:  // plunder.cpp:46-64
        g.bg_group = "water_surface_backgrounds";
        T[7] = {"misc_assets/ship_1.png", "misc_assets/ship_2.png", "misc_assets/ship_3.png", "misc_assets/ship_4.png",
                "misc_assets/ship_5.png", "misc_assets/ship_6.png"};
        T[1] = {"misc_assets/cannonBall.png"};
        T[6] = {"misc_assets/panel_wood.png"};
        T[3] = {"misc_assets/target_red2.png"};
        break;
    case GAME_MINER:  // miner.cpp:38-56
        g.bg_group = "platform_backgrounds";
        T[0] = {"misc_assets/robot_greenDrive1.png"};
        T[1] = {"misc_assets/elementStone007.png"};
        T[2] = {"misc_assets/gemBlue.png"};
        T[6] = {"misc_assets/window.png"};
        T[9] = {"misc_assets/dirt.png"};
        T[10] = {"misc_assets/tile_bricksGrey.png"};
        break;
    case GAME_DODGEBALL:  // dodgeball.cpp:46-88
        g.bg_group = "topdown_backgrounds";
        T[0] = {"misc_assets/character12.png"};
        T[3] = {"misc_assets/ball_soccer1.png"};
        for (int i = 1; i <= 11; i++) T[4].push_back("misc_assets/character" + std::to_string(i) + ".png");
        T[5] = {"misc_assets/blockRed.png"};
        T[6] = {"misc_assets/ball_soccer2.png"};
        T[7] = {"misc_assets/blockGreen.png"};
        T[1] = {"misc_assets/tileStone_slope2.png"};
        T[10] = {"misc_assets/tileStone_slope2.png"};
        for (int i = 1; i <= 9; i++) T[8].push_back("misc_assets/spaceEffect" + std::to_string(i) + ".png");
        break;
    case GAME_STARPILOT:  // starpilot.cpp:55-108
        g.bg_group = "space_backgrounds";
        T[0] = {"misc_assets/playerShip2_blue.png"};
        T[1] = {"misc_assets/towerDefense_tile295.png"};
        T[2] = {"misc_assets/towerDefense_tile296.png"};
        T[3] = {"misc_assets/towerDefense_tile297.png"};
        for (int i = 1; i <= 7; i++) {
            T[4].push_back("misc_assets/spaceShips_00" + std::to_string(i) + ".png");
            T[8].push_back("misc_assets/spaceShips_00" + std::to_string(i) + ".png");
        }
        for (int i = 1; i <= 4; i++) T[5].push_back("misc_assets/spaceMeteors_00" + std::to_string(i) + ".png");
        for (int i = 1; i <= 4; i++) T[5].push_back("misc_assets/meteorGrey_big" + std::to_string(i) + ".png");
        for (int i = 1; i <= 9; i++) T[6].push_back("misc_assets/spaceEffect" + std::to_string(i) + ".png");
        T[7] = {"misc_assets/spaceStation_018.png", "misc_assets/spaceStation_019.png"};
        for (int i = 1; i <= 4; i++) T[9].push_back("misc_assets/spaceRockets_00" + std::to_string(i) + ".png");
        break;
    case GAME_JUMPER:  // jumper.cpp:50-78
        g.bg_group = "platform_backgrounds";
        T[0] = {"misc_assets/bunny2_ready.png"};
        T[2] = {"misc_assets/spikeMan_stand.png"};
        T[1] = {"misc_assets/carrot.png"};
        T[9] = {"misc_assets/bunny2_jump.png"};
        T[12] = {"misc_assets/bunny2_walk1.png"};
        T[13] = {"misc_assets/bunny2_walk2.png"};
        T[10] = {"misc_assets/bunny2_walk1.png"};
        T[11] = {"misc_assets/bunny2_walk2.png"};
        T[7] = {"platformer/tileBlue_05.png", "platformer/tileGreen_05.png", "platformer/tileYellow_06.png", "platformer/tileBrown_06.png"};
        T[6] = {"platformer/tileBlue_08.png", "platformer/tileGreen_08.png", "platformer/tileYellow_09.png", "platformer/tileBrown_09.png"};
        break;
    default:
        throw std::runtime_error("procgen_b200: game id " + std::to_string(game_id) + " has no asset table yet");
    }
    return g;
}

}  // namespace host
}  // namespace pg
