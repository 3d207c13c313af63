// Drop-in for the reference's map.hpp (map.hpp:1-40): same macros and class map::Map.  The certainty
// grid and the lookup table live on the B200; `world` is a host mirror refreshed by syncWorld() (the
// reference's 27 MB array member, map.hpp:25).  pointLookupTable (432 MB in the reference, map.hpp:24) is kept on
// the device as one int per voxel (index of the stored point, -1 = empty); mapCloud receives the same appends.
#ifndef MAP_HPP
#define MAP_HPP

#include "icp.hpp"
#include "pointcloud.hpp"

// The grid is 300^3 cells of 10 m / 300 because the reference fixes it with these macros and a static array member
// (map.hpp:9-10,25: `unsigned char world[MAP_HEIGHT][MAP_HEIGHT][MAP_HEIGHT]`), which cannot express a non-cubic grid.
// The grids of the README / BASELINE configs (300x300x250 at 2 cm, 600x600x500 at 1 cm) and z-slab sharding are a
// run-time choice of the C-ABI underneath: icpb_map_create(ctx, dims, cell, z_lo, z_hi, ...) and icpb_slabmap_*
// (include/icpb200.h; INTEGRATION.md sections B and C).
#define MAP_HEIGHT 300
#define PHYSICAL_HEIGHT 10.0f
#define DELTA_CONFIDENCE 25
#define MIN_CONFIDENCE 50
#define MAX_CONFIDENCE 180
#define MAX_POINT_ADD_DISTANCE 0.05f
#define MAX_KEYPOINT_ADD_DISTANCE 0.1f

#define CELL_PHYSICAL_HEIGHT PHYSICAL_HEIGHT / ((float) MAP_HEIGHT)

struct icpb_map;

namespace map {
class Map;
// pointLookupTable[x][y][z] (map.hpp:24) as a read-only view: the reference's 432 MB array of stored points is an
// int-per-voxel table on the device here; the view fetches the entry and returns the stored point (or `empty`).
struct LookupZ {
    const Map *m; int x, y;
    color_point_t operator[](int z) const;
};
struct LookupY {
    const Map *m; int x;
    LookupZ operator[](int y) const { return LookupZ{m, x, y}; }
};
struct LookupTable {
    const Map *m;
    LookupY operator[](int x) const { return LookupY{m, x}; }
};

class Map { // map.hpp:20-37
public:
    color_point_t empty;
    icp::PointCloud mapCloud;
    LookupTable pointLookupTable;                   // map.hpp:24 (read-only view, see above)
    unsigned char (*world)[MAP_HEIGHT][MAP_HEIGHT]; // world[x][y][z], host mirror of the device grid

    Map();
    ~Map();
    Map(const Map &) = delete;
    Map &operator=(const Map &) = delete;
    void update(icp::PointCloud data, int delta_confidence, cv::viz::Viz3d &depthWindow);                  // map.cpp:220-269
    void update(associations_t associations, int delta_confidencec);                                        // map.cpp:88-119
    void update(associations_t keyPointAssociations, std::vector<float> errors, point_list_t nonAssociations,
                int delta_confidence);                                                                      // map.cpp:122-151
    // map.hpp:31 declares a fourth overload that map.cpp never defines (using it does not link in the reference
    // either); declared here for header parity and left undefined in the same way
    void update(associations_t keyPointAssociations, std::vector<float> errors, icp::PointCloud &dataCloud, int delta_confidence);
    void rayTrace(cv::Point3i point, cv::Point3i origin, cv::viz::Viz3d &depthWindow);                      // map.cpp:272-439
    void drawCertaintyMap(cv::viz::Viz3d &depthWindow);
    cv::Point3i getVoxelCoordinates(cv::Point3f);                                                           // map.cpp:55-85
    bool isOccupied(cv::Point3f);                                                                           // map.cpp:441-444
    int bound(int t, int ds);

    // Additions over the reference
    void integrateRays(icp::PointCloud &cloud, cv::Point3f origin, int delta_dec, int delta_inc); // whole-cloud M4
    void syncWorld();                                                                             // device -> world
    void clear();
    color_point_t lookup(int x, int y, int z) const;                                              // pointLookupTable[x][y][z]

private:
    icpb_map *dev_;
    void ensure();
};
} // namespace map

#endif
