// Drop-in for the reference's icp.hpp (icp.hpp:1-47): same macros, associations_t and namespace icp
// entry points.  Bulk paths (association scans, the registration loop) run on the B200.
#ifndef ICP_HPP
#define ICP_HPP

#define PI 3.14159265358979f
#define MAX_NN_POINT_DISTANCE 1.5f
#define COLOR_WEIGHT 0.0f
#define DISTANCE_WEIGHT 1.0f - COLOR_WEIGHT
#define MAX_NN_COLOR_DISTANCE 0.75f

#define MAX_NN_KEYPOINT_DISTANCE 0.1f
#define MIN_NN_COLOR_DISTANCE 0.2f

#define MIN_ASSOCIATION_DRAW_DISTANCE 0.05f
#define MAX_TRANSLATION_NN_DISTANCE 0.3f
#define MAX_TRANSLATION_DISTANCE 3.5f

#include "pointcloud.hpp"

namespace map { class Map; }

typedef std::vector<std::pair<color_point_t, color_point_t>> associations_t;

namespace icp {
// icp.cpp:28-285 with the all-point association of :149/:253 (the north-star path); whole loop on the device.
cv::Mat getTransformation(cv::Mat &data, cv::Mat &previous, cv::Mat color, std::vector<cv::KeyPoint> keypoints,
                          cv::Mat &rotation, int maxIterations, float threshold, cv::viz::Viz3d &depthWindow);
cv::Mat makeRotationMatrix(float x, float y, float z);
float meanSquareError(std::vector<float> errors);
void showAssocations(associations_t associations, std::vector<float> errors, cv::viz::Viz3d &depthWindow);
cv::Point3f calculateOffset(associations_t associations);
float distance(cv::Point3f a, cv::Point3f b);
float distance(color_point_t a, color_point_t b);
void findGlobalNearestNeighborAssociations(PointCloud &data, PointCloud &previous, std::vector<float> &errors,
                                           associations_t &associations);
void findGlobalKeyPointAssociations(PointCloud &data, std::vector<float> &errors, associations_t &associations,
                                    point_list_t &nonAssociations);
void findMappedNearestNeighborAssociations(PointCloud &data, std::vector<float> &errors, associations_t &associations);
// icp.cpp:476-486 and :371-474, the helpers of the (dead, out-of-bounds-reading) expanding-cube search.  processVoxel
// keeps its meaning -- the lookup-table entry of voxel (x, y, z), if any, against the running best -- as a host scalar
// helper over the map cloud (each voxel's entry is the one map-cloud point recorded in it); getNearestMappedPoint is
// made exact: the nearest point of the map cloud, MAX_NN_COLOR_DISTANCE when nothing is closer (what :379,:474 return).
void processVoxel(color_point_t point, color_point_t &nearest, float &shortestDistance, int x, int y, int z);
float getNearestMappedPoint(color_point_t point, color_point_t &nearest);
float getNearestPoint(color_point_t point, color_point_t &nearest, PointCloud &cloud);
float getNearestKeyPoint(color_point_t point, color_point_t &nearest, PointCloud &cloud);

// Additions over the reference (explicit state instead of icp.cpp:22-26's file-scope globals)
// getTransformation's association: ALL_POINTS (default) is the all-point scan of icp.cpp:149/253 against the previous
// frame (the north-star path); KEYPOINTS is the loop exactly as the reference runs it (icp.cpp:98/255/271): key-points
// against the growing map cloud, rule-C map update from the rejects.
enum { ASSOCIATE_ALL_POINTS = 0, ASSOCIATE_KEYPOINTS = 1 };
void setAssociationMode(int mode);
void resetState();
map::Map &mapState(); // the process-global map of icp.cpp:26
cv::Mat cameraRotationState();
cv::Point3f cameraPositionState();
} // namespace icp

#endif
