// TEST INFRASTRUCTURE ONLY -- the handle type shared by ref_driver.cpp and dropin_driver.cpp.
#ifndef RSM_ORACLE_REF_TYPES_H_
#define RSM_ORACLE_REF_TYPES_H_
#include <memory>
struct RefMap {
  std::shared_ptr<roborts_slam::ScanMatchMap> map;
};
#endif
