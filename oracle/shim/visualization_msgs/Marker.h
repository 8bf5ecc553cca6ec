// oracle/shim/visualization_msgs/Marker.h — TEST INFRASTRUCTURE: plain structs with the fields kino_astar.cpp fills (RViz only)
#pragma once
#include <ros/ros.h>
#include <string>
#include <vector>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Vector3 { double x = 0, y = 0, z = 0; };
}  // namespace geometry_msgs
namespace std_msgs {
struct Header { std::string frame_id; ros::Time stamp; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}  // namespace std_msgs
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, ADD = 0 };
  std_msgs::Header header;
  std::string ns;
  int id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
};
}  // namespace visualization_msgs
