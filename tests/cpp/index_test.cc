// The reference's gtest cases for ANNIndex (embeddinghub/embeddingstore/test/index_test.cc:17-60),
// restated without gtest, against the drop-in twin in include/ehb200_ann_index.hpp.
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "ehb200_ann_index.hpp"

using featureform::embedding::ANNIndex;

static int failures = 0;
static void expect_eq(const char* name, const std::vector<std::string>& got, const std::vector<std::string>& want) {
  if (got != want) {
    ++failures;
    std::printf("FAIL %s: got [", name);
    for (auto& s : got) std::printf("%s ", s.c_str());
    std::printf("]\n");
  } else {
    std::printf("ok   %s\n", name);
  }
}
static std::unique_ptr<ANNIndex> abc() {
  auto idx = std::make_unique<ANNIndex>(3);
  idx->set("a", {0, 1, 0});
  idx->set("b", {1, 1, 0});
  idx->set("c", {1, 0, 0});
  return idx;
}
int main() {
  expect_eq("TestSimpleANN", abc()->approx_nearest({0, 1, 0}, 1), {"a"});
  expect_eq("TestMultiANN", abc()->approx_nearest({0, 1, 0}, 2), {"a", "b"});
  {
    auto idx = abc();
    idx->set("a", {0, -1, 0});
    expect_eq("TestUpdateANN", idx->approx_nearest({0, 1, 0}, 1), {"b"});
  }
  expect_eq("TestANN0Items", abc()->approx_nearest({0, 1, 0}, 0), {});
  return failures ? 1 : 0;
}
