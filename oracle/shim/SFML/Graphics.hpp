// TEST INFRASTRUCTURE ONLY (oracle/): self-written stand-in for the slice of SFML the
// reference's shading code names (utils.hpp:3, raycaster.hpp:4): sf::Color, sf::Vector2<T>
// and an sf::Image with create/loadFromFile/getSize/getPixel/setPixel.
// loadFromFile does not touch the file system: the harness registers 16x16 RGB textures by
// file name (vrt_ref_register_texture) because /root/reference does not exist on the GPU box.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace sf {

typedef std::uint8_t Uint8;

struct Color {
    Uint8 r, g, b, a;
    Color() : r(0), g(0), b(0), a(255) {}
    Color(Uint8 r_, Uint8 g_, Uint8 b_, Uint8 a_ = 255) : r(r_), g(g_), b(b_), a(a_) {}
    static const Color Black, White, Red, Green, Blue, Yellow, Magenta, Cyan, Transparent;
};

template <typename T> struct Vector2 {
    T x, y;
    Vector2() : x(0), y(0) {}
    Vector2(T x_, T y_) : x(x_), y(y_) {}
};
typedef Vector2<int> Vector2i;
typedef Vector2<unsigned int> Vector2u;
typedef Vector2<float> Vector2f;

// name -> (w, h, top-down RGB rows); filled by the harness.
std::map<std::string, std::vector<Uint8>>& shim_texture_registry();

class Image {
public:
    void create(unsigned int w, unsigned int h, const Color& c = Color(0, 0, 0)) {
        m_w = w; m_h = h;
        m_px.assign(size_t(w) * h, c);
    }
    bool loadFromFile(const std::string& path) {
        const size_t cut = path.find_last_of('/');
        const std::string name = cut == std::string::npos ? path : path.substr(cut + 1);
        auto& reg = shim_texture_registry();
        auto it = reg.find(name);
        if (it == reg.end()) return false;
        const std::vector<Uint8>& d = it->second;   // [w, h, rgb...]
        create(d[0], d[1]);
        for (size_t i = 0; i < size_t(m_w) * m_h; ++i) m_px[i] = Color(d[2 + 3 * i], d[3 + 3 * i], d[4 + 3 * i]);
        return true;
    }
    Vector2u getSize() const { return Vector2u(m_w, m_h); }
    Color getPixel(unsigned int x, unsigned int y) const { return m_px[size_t(y) * m_w + x]; }
    void setPixel(unsigned int x, unsigned int y, const Color& c) { m_px[size_t(y) * m_w + x] = c; }
    const Color* data() const { return m_px.data(); }
private:
    unsigned int m_w = 0, m_h = 0;
    std::vector<Color> m_px;
};

}  // namespace sf
