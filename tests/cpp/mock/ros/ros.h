// MOCK of ros::NodeHandle (TEST INFRASTRUCTURE): the hooks of QMController take one by reference.
#pragma once
namespace ros { struct NodeHandle {}; }
