// oracle/shim/ros/ros.h — TEST INFRASTRUCTURE: the few roscpp names kino_astar.cpp uses, without ROS.
// NodeHandle::param reads from a std::map the driver fills; subscribe keeps the member callback so that the driver can deliver
// one "local_cloud" message; publishers swallow their messages.
#pragma once
#include <boost/make_shared.hpp>
#include <algorithm>  // roscpp pulls these in for the reference sources (rrt_star.cpp uses std::queue, assert, std::remove bare)
#include <cassert>
#include <queue>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace ros {
struct Duration { double s; double toSec() const { return s; } };
// the driver of a search whose termination depends on the clock (RRTStar::search, rrt_star.cpp:413-418) installs a deterministic one
inline double (*&time_hook())() { static double (*h)() = nullptr; return h; }
struct Time {
  double t = 0;
  static Time now() { Time x; if (time_hook()) x.t = time_hook()(); return x; }
  Duration operator-(const Time& o) const { return Duration{t - o.t}; }
};
struct Publisher {
  template <typename M> void publish(const M&) const {}
};
struct Subscriber {};
struct NodeHandle {
  std::map<std::string, double> values;                                   // set by the driver
  std::map<std::string, std::function<void(const void*)>> callbacks;      // topic -> type-erased callback
  template <typename T> void param(const std::string& name, T& var, const T& def) {
    auto it = values.find(name);
    var = (it == values.end()) ? def : (T)it->second;
  }
  void param(const std::string& name, double& var, double def) { param<double>(name, var, def); }
  void param(const std::string& name, int& var, int def) { param<int>(name, var, def); }
  template <typename M, typename C>
  Subscriber subscribe(const std::string& topic, int, void (C::*fn)(const boost::shared_ptr<const M>&), C* obj) {
    callbacks[topic] = [fn, obj](const void* p) { (obj->*fn)(*static_cast<const boost::shared_ptr<const M>*>(p)); };
    return Subscriber();
  }
  template <typename M> Publisher advertise(const std::string&, int) { return Publisher(); }
  template <typename M> void deliver(const std::string& topic, const boost::shared_ptr<const M>& msg) {
    callbacks.at(topic)(&msg);
  }
};
}  // namespace ros
