// ref_harness.cpp — drives the REAL reference host code (TEST INFRASTRUCTURE ONLY).
//
// This file contains no reference code.  It #includes the reference's single translation unit
// (RayTracing/src/rayTracing.cpp, with `main` renamed) from where it lies under /root/reference and
// calls the reference's own functions — createClassicCornellBox / addCornellBox /
// addMirrorCornellBox / ... (rayTracing.cpp:388-1118), BVH (BVH.h:145-221), Camera
// (camera.h:99-192), getTrianglesData_ (mesh.h:279-613), random (external/math/random.h:4-10) —
// dumping their results as raw arrays so that tests can pin the oracle and the product's host code
// to them.  Built by oracle/Makefile into oracle/_ref/ref_host (git-ignored).  GLFW / GL / dialog
// symbols stay unresolved at link time (-Wl,--unresolved-symbols=ignore-all): nothing here calls
// them.  The GLSL hot path has its own harness: ref_shader.cpp (the shader source through glsl2cpp.py + glm).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <sstream>

// stb_image's implementation is instantiated here (the reference does it in external/stb/stb_impl.cpp,
// whose stb_image_write half needs MSVC's sprintf_s)
#define STB_IMAGE_IMPLEMENTATION
#define main ref_main
#include "RayTracing/src/rayTracing.cpp"
#undef main
#include <math/random.h>

// portable stand-ins for the two Windows / GL-bound helpers the loader needs
std::vector<std::string> getFilenamesInFolder(const std::string& folderPath) {
    std::vector<std::string> names;
    if (!std::filesystem::exists(folderPath)) return names;
    for (const auto& e : std::filesystem::directory_iterator(folderPath))
        if (e.is_regular_file()) names.push_back(e.path().filename().string());
    std::sort(names.begin(), names.end());
    return names;
}
struct LoadedTex {
    int w, h, ch;
    std::vector<unsigned char> px;
};
static std::vector<LoadedTex> g_textures;
Texture2D::Texture2D(const std::string& path, GLenum textureUnit) {
    unit = textureUnit;
    stbi_set_flip_vertically_on_load(true);  // textureClass.cpp:65
    int w, h, ch;
    unsigned char* p = stbi_load(path.c_str(), &w, &h, &ch, 0);
    if (!p) throw std::runtime_error("Failed to load texture");
    LoadedTex t{w, h, ch, std::vector<unsigned char>(p, p + (size_t)w * h * ch)};
    stbi_image_free(p);
    g_textures.push_back(std::move(t));
}

static_assert(sizeof(RTXTriangle) == 80 && sizeof(Material) == 96 && sizeof(Node) == 48 &&
                  sizeof(GlobalUniforms) == 192,
              "wire layout");

// RTSC container: 8 x int64 header {magic, n_tris, n_mats, n_nodes, n_perm, n_tex, 0, 0}, then
// tris(80 B), mats(96 B), nodes(48 B), permuted tris(80 B), then per texture {i32 w,h,ch,0; bytes}
static const int64_t MAGIC = 0x43535452;  // "RTSC"
struct SceneFile {
    std::vector<RTXTriangle> tris;
    std::vector<Material> mats;
    std::vector<Node> nodes;
    std::vector<RTXTriangle> perm;
    std::vector<LoadedTex> tex;
};
static bool writeScene(const char* path, const SceneFile& s) {
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    int64_t hdr[8] = {MAGIC, (int64_t)s.tris.size(), (int64_t)s.mats.size(), (int64_t)s.nodes.size(),
                      (int64_t)s.perm.size(), (int64_t)s.tex.size(), 0, 0};
    fwrite(hdr, 8, 8, f);
    fwrite(s.tris.data(), 80, s.tris.size(), f);
    fwrite(s.mats.data(), 96, s.mats.size(), f);
    fwrite(s.nodes.data(), 48, s.nodes.size(), f);
    fwrite(s.perm.data(), 80, s.perm.size(), f);
    for (const auto& t : s.tex) {
        int32_t th[4] = {t.w, t.h, t.ch, 0};
        fwrite(th, 4, 4, f);
        fwrite(t.px.data(), 1, t.px.size(), f);
    }
    fclose(f);
    return true;
}
static bool readScene(const char* path, SceneFile& s) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    int64_t hdr[8];
    if (fread(hdr, 8, 8, f) != 8 || hdr[0] != MAGIC) return false;
    RTXTriangle proto(0, glm::vec4(0), glm::vec4(0), glm::vec4(0), glm::vec2(0), glm::vec2(0), glm::vec2(0));
    s.tris.assign((size_t)hdr[1], proto);
    s.mats.resize((size_t)hdr[2]);
    if (fread((void*)s.tris.data(), 80, s.tris.size(), f) != s.tris.size()) return false;
    if (fread((void*)s.mats.data(), 96, s.mats.size(), f) != s.mats.size()) return false;
    fclose(f);
    return true;
}

static void pushFixedMaterials(std::vector<Material>& materials) {  // rayTracing.cpp:1268-1283
    Material red;
    red.makeDiffusive(glm::vec3(1.0f, 0.0f, 0.0f));
    materials.push_back(red);
    Material green;
    green.makeDiffusive(glm::vec3(0.0f, 1.0f, 0.0f));
    materials.push_back(green);
    Material wall;
    wall.makeDiffusive(glm::vec3(1.0f));
    materials.push_back(wall);
    Material light;
    light.makeLight(glm::vec3(1.0f), CORNELL_LIGHT_BRIGHTNESS);
    materials.push_back(light);
    Material mirror;
    mirror.makeSpecular(glm::vec3(1.0f), glm::vec3(1.0f), 1.0f, 1.0f);
    materials.push_back(mirror);
}

// The Material struct leaves several floats uninitialised (mesh.h:47); zero the storage first so
// dumps are reproducible.  Placement-new keeps the reference constructor's own writes.
static Material cleanMaterial() {
    alignas(16) unsigned char buf[sizeof(Material)];
    memset(buf, 0, sizeof buf);
    Material* m = new (buf) Material();
    return *m;
}

static int applyContainer(const std::string& kind, std::vector<RTXTriangle>& rtx,
                          std::vector<BVHTriangle>& bvh, std::vector<Material>& materials) {
    const int n = (int)materials.size();
    if (kind == "none") return 0;
    if (kind == "classic") createClassicCornellBox(rtx, bvh, 10, n - 5, n - 4, n - 3, n - 2);
    else if (kind == "cornell") addCornellBox(rtx, bvh, CORNELL_LIGHT_SIZE, CORNELL_PADDING, n - 2, true);
    else if (kind == "mirror") addMirrorCornellBox(rtx, bvh, CORNELL_LIGHT_SIZE, CORNELL_PADDING, n - 2, n - 1);
    else if (kind == "sidelit") addSideLitCornellBox(rtx, bvh, CORNELL_LIGHT_SIZE, CORNELL_PADDING, n - 2, n - 3, 1);
    else if (kind == "sky") addSkyLightPlane(rtx, bvh, n - 2);
    else return 1;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr,
                "usage: ref_host scene <kind> <in.rtsc|-> <out.rtsc>   kind: none classic cornell mirror sidelit sky\n"
                "       ref_host load <model folder> <kind> <out.rtsc>\n"
                "       ref_host camera W H px py pz hfov pitch yaw focus defocus zoom <out.bin>\n"
                "       ref_host rng <seed> <n> <out.bin>\n");
        return 2;
    }
    const std::string cmd = argv[1];
    // silence the reference's chatter (BVH.h:213-219 prints every leaf)
    std::ostringstream sink;
    std::streambuf* oldCout = std::cout.rdbuf(sink.rdbuf());
    int rc = 0;
    try {
        if (cmd == "scene" && argc == 5) {
            SceneFile in, out;
            std::vector<RTXTriangle> rtx;
            std::vector<BVHTriangle> bvh;
            std::vector<Material> materials;
            if (std::string(argv[3]) != "-") {
                if (!readScene(argv[3], in)) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
                rtx = in.tris;
                materials = in.mats;
                for (const auto& t : rtx) bvh.push_back(BVHTriangle(glm::vec3(t.a), glm::vec3(t.b), glm::vec3(t.c)));
            } else {
                Material def = cleanMaterial();  // mesh.h:322-324
                def.index = 0;
                materials.push_back(def);
            }
            {
                std::vector<Material> fixed;
                pushFixedMaterials(fixed);
                for (auto& m : fixed) {  // deterministic padding bytes
                    Material c = cleanMaterial();
                    c.color = m.color; c.materialType = m.materialType;
                    if (m.materialType == LIGHT) { c.emissionColor = m.emissionColor; c.emissionStrength = m.emissionStrength; }
                    if (m.materialType == SPECULAR) { c.specularColor = m.specularColor; c.smoothness = m.smoothness; c.specularProbability = m.specularProbability; }
                    materials.push_back(c);
                }
            }
            if (applyContainer(argv[2], rtx, bvh, materials)) { fprintf(stderr, "bad kind\n"); return 1; }
            // BVH bounds from the triangle's own vertices (works around rayTracing.cpp:535-537)
            bvh.clear();
            for (const auto& t : rtx) bvh.push_back(BVHTriangle(glm::vec3(t.a), glm::vec3(t.b), glm::vec3(t.c)));
            out.tris = rtx;
            out.mats = materials;
            BVH tree(bvh, rtx);  // rayTracing.cpp:1293
            out.nodes = tree.allNodes;
            out.perm = rtx;
            if (!writeScene(argv[4], out)) rc = 1;
        } else if (cmd == "load" && argc == 5) {
            SceneFile out;
            std::vector<RTXTriangle> rtx;
            std::vector<BVHTriangle> bvh;
            std::vector<Material> materials;
            std::vector<Texture2D> textures;
            getTrianglesData_(argv[2], 1, rtx, bvh, materials, textures);  // rayTracing.cpp:1265
            pushFixedMaterials(materials);
            if (applyContainer(argv[3], rtx, bvh, materials)) { fprintf(stderr, "bad kind\n"); return 1; }
            bvh.clear();
            for (const auto& t : rtx) bvh.push_back(BVHTriangle(glm::vec3(t.a), glm::vec3(t.b), glm::vec3(t.c)));
            out.tris = rtx;
            out.mats = materials;
            BVH tree(bvh, rtx);
            out.nodes = tree.allNodes;
            out.perm = rtx;
            out.tex = g_textures;
            if (!writeScene(argv[4], out)) rc = 1;
        } else if (cmd == "camera" && argc == 14) {
            const int W = atoi(argv[2]), H = atoi(argv[3]);
            glm::vec3 pos((float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6]));
            Camera camera(W, H, maxSpeed, pos, (float)atof(argv[7]), (float)atof(argv[8]), (float)atof(argv[9]),
                          (float)atof(argv[10]), (float)atof(argv[11]), (float)atof(argv[12]));
            GlobalUniforms u;
            memset(&u, 0, sizeof u);
            camera.updateUniforms(u);
            u.width = W;
            u.height = H;
            FILE* f = fopen(argv[13], "wb");
            fwrite(&u, sizeof u, 1, f);
            fclose(f);
        } else if (cmd == "rng" && argc == 5) {
            unsigned int state = (unsigned int)strtoul(argv[2], nullptr, 0);
            const int n = atoi(argv[3]);
            FILE* f = fopen(argv[4], "wb");
            for (int i = 0; i < n; i++) {
                float v = random(state);
                fwrite(&state, 4, 1, f);
                fwrite(&v, 4, 1, f);
            }
            fclose(f);
        } else {
            fprintf(stderr, "bad command\n");
            rc = 2;
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        rc = 1;
    } catch (int e) {
        fprintf(stderr, "reference loader threw errno %d\n", e);
        rc = 1;
    }
    std::cout.rdbuf(oldCout);
    return rc;
}
