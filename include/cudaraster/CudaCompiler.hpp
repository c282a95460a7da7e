// FW::CudaCompiler -- run-time compilation of pixel-pipe source files, kept from the reference
// (src/framework/gpu/CudaCompiler.hpp:40-232, gpu/CudaCompiler.cpp:83-193, :417-584): the caller
// names a .cu source that instantiates CR_DEFINE_PIXEL_PIPE, adds -I paths and -D defines
// (SAMPLES_LOG2, RENDER_MODE_FLAGS, BLEND_SHADER ... exactly like test/SceneCR.cpp:170-179), and
// compile() returns a CudaModule the rasterizer resolves the pipe from by name.
//
// Like the reference it shells out to nvcc and caches the result on disk, keyed on everything that
// can change the binary: nvcc version, options, defines, preamble, the source file's size and
// modification time, and those of the pipeline headers it includes.  What is different:
//   * the product of a compile is a SHARED OBJECT for sm_100a (the B200 pipeline's stage entry points
//     are host launchers, see cuda/PixelPipe.inl), not a cubin loaded with cuModuleLoadData;
//   * the memory cache holds dlopen() handles; modules stay owned by the cache
//     (gpu/CudaCompiler.cpp:89-108) until flushMemCache() / staticDeinit().
// Header only; host code; needs nvcc at run time (CUDA_BIN_PATH / CUDA_HOME / PATH).
#pragma once
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "CudaRaster.hpp"

namespace FW {

class CudaCompiler {
public:
    CudaCompiler(void) : m_cachePath("cudacache"), m_overriddenSMArch(0) {}
    ~CudaCompiler(void) {}

    void setCachePath(const std::string& path) { m_cachePath = path; }
    void setSourceFile(const std::string& path) { m_sourceFile = path; }
    void overrideSMArch(int arch) { m_overriddenSMArch = arch; }   // accepted; the pipeline is sm_100a only

    void clearOptions(void) { m_options = ""; }
    void addOptions(const std::string& options) { m_options += options + " "; }
    void include(const std::string& path) { addOptions("-I\"" + path + "\""); }

    void clearDefines(void) { m_defines.clear(); }
    void undef(const std::string& key) { m_defines.erase(key); }
    void define(const std::string& key, const std::string& value = "") { m_defines[key] = value; }
    void define(const std::string& key, int value) { char n[16]; snprintf(n, sizeof(n), "%d", value); define(key, std::string(n)); }

    void clearPreamble(void) { m_preamble = ""; }
    void addPreamble(const std::string& preamble) { m_preamble += preamble + "\n"; }

    // Compiles (or fetches from the caches) and loads the module.  fail()s with the compiler log on error.
    CudaModule* compile(bool enablePrints = true) {
        const std::string file = compileSharedObjectFile(enablePrints);
        if (file.empty()) return NULL;
        std::map<std::string, CudaModule*>& cache = moduleCache();
        std::map<std::string, CudaModule*>::iterator it = cache.find(file);
        if (it != cache.end()) return it->second;
        CudaModule* m = new CudaModule(file);
        cache[file] = m;
        return m;
    }

    // The analogue of compileCubinFile(): path of the built shared object ("" never; errors fail()).
    std::string compileSharedObjectFile(bool enablePrints = true) {
        if (m_sourceFile.empty()) fail("CudaCompiler: No source file specified!");
        if (!fileExists(m_sourceFile)) fail("CudaCompiler: Source file '%s' not found!", m_sourceFile.c_str());
        const std::string nvcc = findNvcc();
        const std::string incRoot = includeRoot();
        const std::string libDir = libraryDir();

        // ---- cache key
        U64 h = 1469598103934665603ull;
        hashStr(h, nvccVersion(nvcc));
        hashStr(h, staticOptions() + "|" + m_options + "|" + m_preamble + "|" + m_sourceFile);
        for (std::map<std::string, std::string>::const_iterator it = m_defines.begin(); it != m_defines.end(); ++it) hashStr(h, it->first + "=" + it->second + ";");
        hashStr(h, fileStamp(m_sourceFile));
        const char* hdrs[] = {"/crb200.h", "/cudaraster/cuda/PixelPipe.inl", "/cudaraster/cuda/PixelPipe.hpp", "/cudaraster/cuda/FineRaster.cuh", "/cudaraster/cuda/FineRasterMSAA.cuh",
                              "/cudaraster/cuda/TriangleSetup.cuh", "/cudaraster/cuda/Overlap.cuh", "/cudaraster/cuda/Util.cuh", "/cudaraster/cuda/PrivateDefs.hpp", "/cudaraster/cuda/Constants.hpp"};
        for (size_t i = 0; i < sizeof(hdrs) / sizeof(hdrs[0]); i++) hashStr(h, fileStamp(incRoot + hdrs[i]));
        char name[64];
        snprintf(name, sizeof(name), "%016llx.so", (unsigned long long)h);
        mkdir(m_cachePath.c_str(), 0777);
        const std::string out = m_cachePath + "/" + name;
        if (fileExists(out)) {
            if (enablePrints) printf("CudaCompiler: '%s' -> cached %s\n", m_sourceFile.c_str(), out.c_str());
            return out;
        }

        // ---- compile
        std::string pre;
        if (!m_preamble.empty() || !staticPreamble().empty()) {
            pre = m_cachePath + "/" + std::string(name) + ".preamble.h";
            FILE* fp = fopen(pre.c_str(), "w");
            if (!fp) fail("CudaCompiler: Cannot write '%s'!", pre.c_str());
            fputs(staticPreamble().c_str(), fp);
            fputs(m_preamble.c_str(), fp);
            fclose(fp);
        }
        char pid[32];
        snprintf(pid, sizeof(pid), ".%d.tmp", (int)getpid());
        const std::string tmp = out + pid, log = out + ".log";
        std::string cmd = "\"" + nvcc + "\" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcudafe --diag_suppress=177 -shared -Xlinker -Bsymbolic";
        cmd += " -I\"" + incRoot + "\" " + staticOptions() + " " + m_options;
        for (std::map<std::string, std::string>::const_iterator it = m_defines.begin(); it != m_defines.end(); ++it)
            cmd += " -D" + it->first + (it->second.empty() ? "" : "=" + it->second);
        if (!pre.empty()) cmd += " -include \"" + pre + "\"";
        cmd += " -o \"" + tmp + "\" \"" + m_sourceFile + "\"";
        if (!libDir.empty()) cmd += " -L\"" + libDir + "\" -lcrb200 -Xlinker -rpath -Xlinker \"" + libDir + "\"";
        cmd += " > \"" + log + "\" 2>&1";
        if (enablePrints) { printf("CudaCompiler: Compiling '%s'...", m_sourceFile.c_str()); fflush(stdout); }
        const int rc = system(cmd.c_str());
        if (rc != 0 || !fileExists(tmp)) {
            std::string text = readFile(log);
            if (text.size() > 4000) text = text.substr(text.size() - 4000);
            fail("CudaCompiler: Compilation of '%s' failed!\n%s\n%s", m_sourceFile.c_str(), cmd.c_str(), text.c_str());
        }
        if (rename(tmp.c_str(), out.c_str()) != 0) fail("CudaCompiler: Cannot write '%s'!", out.c_str());
        if (enablePrints) printf(" Done.\n");
        return out;
    }

    static void setStaticCudaBinPath(const std::string& path) { staticCudaBinPath() = path; }
    static void setStaticOptions(const std::string& options) { staticOptions() = options; }
    static void setStaticPreamble(const std::string& preamble) { staticPreamble() = preamble; }
    static void staticInit(void) {}
    static void staticDeinit(void) { flushMemCache(); }
    static void flushMemCache(void) {
        std::map<std::string, CudaModule*>& cache = moduleCache();
        for (std::map<std::string, CudaModule*>::iterator it = cache.begin(); it != cache.end(); ++it) delete it->second;
        cache.clear();
    }

private:
    CudaCompiler(const CudaCompiler&);             // forbidden
    CudaCompiler& operator=(const CudaCompiler&);  // forbidden

    static std::map<std::string, CudaModule*>& moduleCache(void) { static std::map<std::string, CudaModule*> s; return s; }
    static std::string& staticCudaBinPath(void) { static std::string s; return s; }
    static std::string& staticOptions(void) { static std::string s; return s; }
    static std::string& staticPreamble(void) { static std::string s; return s; }

    static bool fileExists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
    static std::string fileStamp(const std::string& p) {
        struct stat st;
        if (stat(p.c_str(), &st) != 0) return p + ":missing";
        char b[96];
        snprintf(b, sizeof(b), ":%lld:%lld.%09ld", (long long)st.st_size, (long long)st.st_mtim.tv_sec, (long)st.st_mtim.tv_nsec);
        return p + b;
    }
    static void hashStr(U64& h, const std::string& s) {   // FNV-1a
        for (size_t i = 0; i < s.size(); i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
        h ^= 0xFF; h *= 1099511628211ull;
    }
    static std::string readFile(const std::string& p) {
        std::string r;
        FILE* fp = fopen(p.c_str(), "rb");
        if (!fp) return r;
        char buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) r.append(buf, n);
        fclose(fp);
        return r;
    }
    static std::string runCapture(const std::string& cmd) {
        std::string r;
        FILE* fp = popen(cmd.c_str(), "r");
        if (!fp) return r;
        char buf[512];
        while (fgets(buf, sizeof(buf), fp)) r += buf;
        pclose(fp);
        return r;
    }
    // nvcc: setStaticCudaBinPath, then CUDA_BIN_PATH / CUDA_HOME / CUDA_PATH, then /usr/local/cuda, then PATH
    static std::string findNvcc(void) {
        std::vector<std::string> cand;
        if (!staticCudaBinPath().empty()) cand.push_back(staticCudaBinPath() + "/nvcc");
        const char* envs[] = {"CUDA_BIN_PATH", "CUDA_HOME", "CUDA_PATH"};
        for (int i = 0; i < 3; i++) {
            const char* v = getenv(envs[i]);
            if (v && *v) cand.push_back(std::string(v) + (i == 0 ? "/nvcc" : "/bin/nvcc"));
        }
        cand.push_back("/usr/local/cuda/bin/nvcc");
        for (size_t i = 0; i < cand.size(); i++)
            if (fileExists(cand[i])) return cand[i];
        std::string w = runCapture("command -v nvcc 2>/dev/null");
        while (!w.empty() && (w[w.size() - 1] == '\n' || w[w.size() - 1] == ' ')) w.erase(w.size() - 1);
        if (w.empty()) fail("CudaCompiler: Unable to detect CUDA Toolkit binary path!\nPlease set CUDA_BIN_PATH environment variable.");
        return w;
    }
    static std::string nvccVersion(const std::string& nvcc) {
        static std::map<std::string, std::string> s;
        std::map<std::string, std::string>::iterator it = s.find(nvcc);
        if (it != s.end()) return it->second;
        return s[nvcc] = runCapture("\"" + nvcc + "\" --version 2>&1");
    }
    // The pipeline's include root: CRB200_INCLUDE, else the location of this header when the application was
    // built (if that was an absolute path), else <dir of libcrb200.so>/../include (the repository layout).
    static std::string includeRoot(void) {
        std::vector<std::string> cand;
        const char* env = getenv("CRB200_INCLUDE");
        if (env && *env) cand.push_back(env);
        std::string p(__FILE__);
        for (int up = 0; up < 2; up++) {
            const size_t k = p.find_last_of('/');
            p = k == std::string::npos ? std::string(".") : p.substr(0, k);
        }
        if (!p.empty() && p[0] == '/') cand.push_back(p);
        const std::string lib = libraryDir();
        if (!lib.empty()) cand.push_back(lib + "/../include");
        cand.push_back(p);
        for (size_t i = 0; i < cand.size(); i++)
            if (fileExists(cand[i] + "/cudaraster/cuda/PixelPipe.inl")) return cand[i];
        fail("CudaCompiler: cannot find <cudaraster/cuda/PixelPipe.inl>; set CRB200_INCLUDE to the pipeline's include directory!");
        return p;
    }
    // directory of the libcrb200.so the application is linked with
    static std::string libraryDir(void) {
        Dl_info info;
        if (!dladdr((const void*)&crb_abi_version, &info) || !info.dli_fname) return "";
        std::string p(info.dli_fname);
        const size_t k = p.find_last_of('/');
        return k == std::string::npos ? std::string(".") : p.substr(0, k);
    }

    std::string m_cachePath, m_sourceFile, m_options, m_preamble;
    std::map<std::string, std::string> m_defines;
    S32 m_overriddenSMArch;
};

}  // namespace FW
