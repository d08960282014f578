// cvx_last_error() plumbing shared by the translation units that implement include/cvx.h.
#ifndef CVX_API_ERROR_H_
#define CVX_API_ERROR_H_
#include <string>
namespace cvx {
// Records `msg` as the calling thread's last error and returns `code`.
int SetApiError(int code, const std::string& msg);
}  // namespace cvx
#endif  // CVX_API_ERROR_H_
