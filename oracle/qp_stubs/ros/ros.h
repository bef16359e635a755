// oracle/qp_stubs/ros/ros.h -- stand-in for <ros/ros.h>: just enough for planner/qp_solver.hpp to compile VERBATIM
// (QPConfig reads three parameters from a NodeHandle).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <map>
#include <string>
namespace ros {
class NodeHandle {
public:
    std::map<std::string, double> values;
    bool getParam(const std::string &key, double &out) const {
        auto it = values.find(key);
        if (it == values.end()) return false;
        out = it->second;
        return true;
    }
    bool getParam(const std::string &key, int &out) const {
        auto it = values.find(key);
        if (it == values.end()) return false;
        out = static_cast<int>(it->second);
        return true;
    }
};
}  // namespace ros
