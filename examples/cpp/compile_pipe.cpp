// compile_pipe <source.cu> <cacheDir> [KEY=VALUE ...]
// Compiles a pixel-pipe source with FW::CudaCompiler (no GPU needed: nvcc cross-compiles) and prints
// the path of the cached shared object -- the offline twin of what an application does at run time.
#include <cudaraster/CudaCompiler.hpp>

#include <cstring>

int main(int argc, char** argv) {
    if (argc < 3) { printf("usage: compile_pipe <source.cu> <cacheDir> [KEY=VALUE ...]\n"); return 2; }
    FW::CudaCompiler c;
    c.setSourceFile(argv[1]);
    c.setCachePath(argv[2]);
    for (int i = 3; i < argc; i++) {
        const char* eq = strchr(argv[i], '=');
        if (eq) c.define(std::string(argv[i], eq - argv[i]), std::string(eq + 1));
        else c.define(argv[i]);
    }
    const std::string so = c.compileSharedObjectFile();
    printf("%s\n", so.c_str());
    return so.empty() ? 1 : 0;
}
