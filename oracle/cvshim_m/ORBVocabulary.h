// ORACLE (test infrastructure).  Stand-in for include/ORBVocabulary.h (DBoW2 TemplatedVocabulary needs cv::FileStorage):
// the matcher translation unit only passes ORBVocabulary pointers around.  The real header chain (TemplatedVocabulary.h)
// leaks `using namespace std;` into every reference header included after it, and those headers rely on it.
#ifndef ORBVOCABULARY_H
#define ORBVOCABULARY_H
#include <list>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>
using namespace std;
namespace ORB_SLAM2 { class ORBVocabulary {}; }
#endif
